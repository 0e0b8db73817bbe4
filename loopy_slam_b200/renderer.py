"""Host-side mirror of ``Renderer`` (/root/reference/src/utils/Renderer.py:6-276): same constructor
and method signatures, so ``src/Mapper.py`` / ``src/Tracker.py`` call it unchanged.  Every method
drives the fused sm_100a kernels through the C ABI (include/lsr.h); nothing here computes on the
CPU and there is no PyTorch-eager fallback.
"""
import ctypes
import os
import warnings

import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr

DEFAULT_MAX_CELLS = 1 << 22


# ----------------------------------------------------------------------------- neighbour index
CELL_PER_RADIUS = float(os.environ.get('LSR_CELL_FACTOR', '0.5'))   # grid cell edge / query radius (tuning knob)


class GridIndex:
    """Device workspace of one uniform-grid neighbour index (lsr_grid_build)."""

    def __init__(self, cloud_pos, cell, max_cells=DEFAULT_MAX_CELLS):
        """cell: the largest query radius this index will serve; the grid's cell edge is CELL_PER_RADIUS x
        that (the walk prunes cell rows against the query ball, so finer cells trim the candidate list)."""
        _lib.require_cuda(cloud_pos, 'cloud_pos')
        cloud = cloud_pos.detach()
        if cloud.dtype != torch.float32 or not cloud.is_contiguous():
            cloud = cloud.to(torch.float32).contiguous()
        self.cloud = cloud.reshape(-1, 3)
        self.n = self.cloud.shape[0]
        self.cell = float(cell) * CELL_PER_RADIUS
        self.max_cells = int(max_cells)
        nbytes = ctypes.c_size_t()
        check(lib().lsr_grid_workspace_bytes(self.n, self.max_cells, ctypes.byref(nbytes)), 'lsr_grid_workspace_bytes')
        self.ws = torch.empty(nbytes.value, dtype=torch.uint8, device=cloud.device)
        with _lib.on_device(cloud.device):
            check(lib().lsr_grid_build(ptr(self.cloud), self.n, self.cell, self.max_cells, ptr(self.ws), nbytes.value,
                                       stream_ptr(cloud.device)), 'lsr_grid_build')

    def query(self, pos, radius, dynamic_radius=None):
        """== find_neighbors_faiss: D (P,8) f32, I (P,8) i64, neighbor_num (P,) i32."""
        _lib.require_cuda(pos, 'pos')
        q = pos.detach().to(torch.float32).reshape(-1, 3).contiguous()
        P = q.shape[0]
        D = torch.empty(P, 8, dtype=torch.float32, device=q.device)
        I = torch.empty(P, 8, dtype=torch.int64, device=q.device)
        n = torch.empty(P, dtype=torch.int32, device=q.device)
        rd = None
        if dynamic_radius is not None:
            rd = dynamic_radius.detach().to(torch.float64).reshape(-1).contiguous()
            assert rd.shape[0] == P, 'shape mis-match for input points and dynamic radius'
        with _lib.on_device(q.device):
            check(lib().lsr_knn_query(ptr(self.ws), ptr(q), ptr(rd), float(radius), P, ptr(D), ptr(I), ptr(n),
                                      stream_ptr(q.device)), 'lsr_knn_query')
        return D, I, n


class _GridCache:
    """Grid per (cloud tensor identity, version, cell): both callers pass ``cloud_pos`` on every
    render call (src/Tracker.py:166, src/Mapper.py:686); rebuild only when it changed."""

    def __init__(self):
        self.key = None
        self.grid = None

    def get(self, cloud_pos, cell):
        key = (cloud_pos.data_ptr(), cloud_pos.shape[0], cloud_pos._version, float(cell), str(cloud_pos.device))
        if key != self.key:
            self.grid = GridIndex(cloud_pos, cell)
            self.key = key
        return self.grid


def _grid_for(npc, cloud_pos, cell, cache):
    g = getattr(npc, 'grid_index', None)
    if callable(g):          # our NeuralPointCloud keeps its own index in sync with its cloud
        grid = g()
        if grid is not None and (cloud_pos is None or cloud_pos.shape[0] == grid.n):
            return grid
    if cloud_pos is None:
        raise RuntimeError('render needs cloud_pos (or an npc exposing grid_index())')
    return cache.get(cloud_pos, cell)


_RELPOS_FIELDS = ('c_Brel', 'c_nb1_w', 'c_nb1_b', 'c_nb2_w', 'c_nb2_b')


# ----------------------------------------------------------------------------- fused render op
class _RenderCtx:
    """Per-call constants shared by forward and backward."""
    __slots__ = ('prm', 'grid', 'stage', 'is_tracker', 'blob', 'wstruct', 'flat', 'saved', 'scratch', 'R',
                 'r_query', 'device', 'far_group', 'force_save', 'timing', 'remap', 'renderer', 'z_zero', 'grad_enabled')


def _tick(timing):
    """Optional CUDA-event bracket around one C-ABI launch (bench.py: Renderer._timing = {...})."""
    if timing is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _tock(timing, key, start):
    if timing is None:
        return
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    timing[key].append((start, e))


def _f32c(t):
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    return t


class FeatureSubset:
    """The trainable sub-block of the feature tables (src/Mapper.py:498-505): `indices` are the table rows
    whose features live in the compact leaf tensors passed with it.  Holds the (N,) int32 row->leaf-row map
    the kernels read (built once per frame, when the reference builds `indices`).  Passing
    `feat_subset=(subset, geo_leaf, col_leaf)` to Renderer.render_batch_ray replaces the per-iteration
    `npc_geo_feats[indices] = geo_pcl_grad` / `npc_col_feats[indices] = color_pcl_grad` index_puts
    (src/Mapper.py:581-582) and the table-sized gradients + gathers of their backward."""

    def __init__(self, indices, n_points, device=None):
        idx = torch.as_tensor(indices, dtype=torch.int64, device=device).reshape(-1)
        if not idx.is_cuda:
            raise RuntimeError('FeatureSubset: indices must live on (or be sent to) a CUDA device')
        self.indices = idx
        self.n_points = int(n_points)
        self.remap = torch.full((self.n_points,), -1, dtype=torch.int32, device=idx.device)
        self.remap[idx] = torch.arange(idx.numel(), dtype=torch.int32, device=idx.device)

    def __len__(self):
        return int(self.indices.numel())


class _RenderFn(torch.autograd.Function):
    """depth, var, rgb, valid = lsr_render_fwd(...); backward = lsr_render_bwd(...)."""

    @staticmethod
    def forward(ctx, rc, rays_o, rays_d, gt_depth, geo_feats, col_feats, affine, far_zero, geo_leaf, col_leaf,
                *params):
        ctx.set_materialize_grads(False)
        dev = rays_o.device
        R = rays_o.shape[0]
        depth = torch.empty(R, dtype=torch.float32, device=dev)
        var = torch.empty(R, dtype=torch.float32, device=dev)
        rgb = torch.empty(R, 3, dtype=torch.float32, device=dev)
        valid = torch.empty(R, dtype=torch.uint8, device=dev)
        # needs_input_grad stays True under torch.no_grad() (and grad mode is always off INSIDE Function.forward, so the caller's
        # mode is captured in fused_render): without this test every no_grad render (render_img: 816 k rays)
        # would allocate and write the multi-KB-per-sample saved activations
        need_bwd = (rc.grad_enabled and any(ctx.needs_input_grad)) or rc.force_save
        if not need_bwd or (rc.prm.flags & _lib.FLAG_SAVE_LIGHT):
            rc.prm.flags |= _lib.FLAG_FWD_ONLY      # scratch without the backward's hand-over planes
        sb, cb = ctypes.c_size_t(), ctypes.c_size_t()
        check(lib().lsr_render_workspace_bytes(ctypes.byref(rc.prm), R, rc.stage, ctypes.byref(sb), ctypes.byref(cb)),
              'lsr_render_workspace_bytes')
        rc.scratch = torch.empty(cb.value, dtype=torch.uint8, device=dev)
        rc.saved = torch.empty(sb.value, dtype=torch.uint8, device=dev) if need_bwd else None
        ev = _tick(rc.timing)
        with _lib.on_device(dev):      # the kernels launch on dev's stream: make it the current device
            check(lib().lsr_render_fwd(ctypes.byref(rc.prm), ptr(rc.grid.ws), ptr(rc.grid.cloud), rc.grid.n,
                                       ptr(rays_o), ptr(rays_d), ptr(gt_depth), ptr(rc.r_query), ptr(far_zero),
                                       rc.far_group, ptr(rc.z_zero), R, ptr(geo_feats), ptr(col_feats), ptr(rc.remap), ptr(geo_leaf),
                                       ptr(col_leaf), ctypes.byref(rc.wstruct),
                                       ptr(affine), rc.stage, ptr(depth), ptr(var), ptr(rgb), ptr(valid),
                                       ptr(rc.saved), ptr(rc.scratch), stream_ptr(dev)), 'lsr_render_fwd')
        _tock(rc.timing, 'fwd', ev)
        ctx.rc = rc
        if rc.blob.shapes is None:      # shapes / contiguous strides of the parameters, once per blob
            rc.blob.shapes = [tuple(p.shape) for p in params]
            rc.blob.strides = [tuple(p.stride()) for p in params]
        ctx.save_for_backward(rays_o, rays_d, gt_depth, geo_feats, col_feats, affine, geo_leaf, col_leaf)
        ctx.mark_non_differentiable(valid)
        return depth, var, rgb, valid

    @staticmethod
    def backward(ctx, g_depth, g_var, g_rgb, _g_valid):
        rc = ctx.rc
        rays_o, rays_d, gt_depth, geo_feats, col_feats, affine, geo_leaf, col_leaf = ctx.saved_tensors
        dev = rays_o.device
        R = rays_o.shape[0]
        need = ctx.needs_input_grad   # (rc, o, d, gt, geo, col, affine, far, geo_leaf, col_leaf, *params)
        sub = rc.remap is not None    # feature gradients go to the leaf blocks, the tables are constants
        blob = rc.blob
        flags = 0
        d_o = d_d = d_geo = d_col = d_aff = d_w = None
        if need[1] or need[2]:
            flags |= _lib.GRAD_RAYS
            d_o = torch.empty(R, 3, dtype=torch.float32, device=dev)
            d_d = torch.empty(R, 3, dtype=torch.float32, device=dev)
        # every accumulated gradient sink lives in ONE zero-filled buffer (one fill launch instead of four)
        want_geo = need[8] if sub else need[4]
        want_col = (need[9] if sub else need[5]) and rc.stage == 1
        want_aff = need[6] and affine is not None
        pneed = need[10:]
        geo_like = geo_leaf if sub else geo_feats
        col_like = col_leaf if sub else col_feats
        sizes = [geo_like.numel() if want_geo else 0, col_like.numel() if want_col else 0,
                 blob.n_elems if any(pneed) else 0, 12 if want_aff else 0]
        sizes = [(n + 3) // 4 * 4 for n in sizes]     # keep every view 16-byte aligned
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev) if sum(sizes) else None
        if rc.renderer is not None:
            # parallel.GradAllReducer can exchange this ONE buffer instead of every parameter's .grad
            rc.renderer.last_grad_buffer = flat
        o0, o1, o2 = sizes[0], sizes[0] + sizes[1], sizes[0] + sizes[1] + sizes[2]
        if want_geo:
            flags |= _lib.GRAD_GEO_FEATS
            d_geo = flat[:geo_like.numel()].view(geo_like.shape)
        if want_col:
            flags |= _lib.GRAD_COL_FEATS
            d_col = flat[o0:o0 + col_like.numel()].view(col_like.shape)
        if want_aff:
            flags |= _lib.GRAD_AFFINE
            d_aff = flat[o2:o2 + 12]
        if any(pneed):
            d_w = flat[o1:o1 + blob.n_elems]
            if any(pneed[k] for k in blob.geo_w_idx):
                flags |= _lib.GRAD_GEO_W
            if any(pneed[k] for k in blob.geo_b_idx):
                flags |= _lib.GRAD_GEO_B
            if rc.stage == 1 and any(pneed[k] for k in blob.col_w_idx):
                flags |= _lib.GRAD_COL_W
        g_depth = _f32c(g_depth) if g_depth is not None else torch.zeros(R, dtype=torch.float32, device=dev)
        g_var = _f32c(g_var) if g_var is not None else None
        g_rgb = _f32c(g_rgb) if g_rgb is not None else None
        ev = _tick(rc.timing)
        with _lib.on_device(dev):      # the kernels launch on dev's stream: make it the current device
            check(lib().lsr_render_bwd(ctypes.byref(rc.prm), ptr(rc.grid.ws), ptr(rc.grid.cloud), rc.grid.n,
                                       ptr(rays_o), ptr(rays_d), ptr(gt_depth), ptr(rc.r_query), R, ptr(geo_feats),
                                       ptr(col_feats), ptr(rc.remap), ptr(geo_leaf), ptr(col_leaf),
                                       ctypes.byref(rc.wstruct), ptr(affine), rc.stage,
                                       1 if rc.is_tracker else 0, ptr(rc.saved), ptr(rc.scratch), ptr(g_depth),
                                       ptr(g_var), ptr(g_rgb), flags, ptr(d_geo), ptr(d_col), ptr(d_w), ptr(d_aff),
                                       ptr(d_o), ptr(d_d), stream_ptr(dev)), 'lsr_render_bwd')
        _tock(rc.timing, 'bwd', ev)
        rc.saved = None
        # Parameters the rendered stage does not touch get NO gradient (None), exactly like autograd on the reference's
        # graph: a zero tensor instead would make the caller's Adam create state for them (src/Mapper.py:524-541 puts the
        # whole colour decoder into the optimiser while 40 % of the iterations render stage 'geometry'), and its bias
        # correction would then be off when the colour stage starts (caught by tests/test_gpu_trajectory.py).
        relpos = bool(rc.prm.flags & _lib.FLAG_REL_POS)
        pgrads = [None] * len(blob.offsets)
        if d_w is not None:
            base = d_w.storage_offset()
            color = rc.stage == 1
            for k in range(len(pgrads)):
                if not pneed[k]:
                    continue
                field = blob.entries[k][0]
                if field[0] != 'g' and not (color and (relpos or field not in _RELPOS_FIELDS)):
                    continue
                pgrads[k] = d_w.as_strided(blob.shapes[k], blob.strides[k], base + blob.offsets[k])   # one op per view
        if sub:
            return (None, d_o if need[1] else None, d_d if need[2] else None, None, None, None, d_aff, None,
                    d_geo, d_col, *pgrads)
        return (None, d_o if need[1] else None, d_d if need[2] else None, None, d_geo, d_col, d_aff, None,
                None, None, *pgrads)


def _make_params(renderer, decoders, stage, coef):
    f = decoders.cfg_flags
    P = _lib.LsrParams()
    P.n_surface = renderer.N_surface
    P.nn_num = f['nn_num']
    P.min_nn_num = f['min_nn_num']
    P.c_dim = 32
    P.near_end_surface = renderer.near_end_surface
    P.far_end_surface = renderer.far_end_surface
    P.near_end = renderer.near_end
    P.sigmoid_coef = coef
    P.radius_query = 0.0
    flags = 0
    if f['encode_rel_pos_in_col']:
        flags |= _lib.FLAG_REL_POS
    if renderer.use_dynamic_radius:
        flags |= _lib.FLAG_DYNAMIC_R
    if renderer.skip_zero_depth_pixel:
        flags |= _lib.FLAG_SKIP_ZERO_DEPTH
    P.flags = flags
    P.rgb_mode = _lib.RGB_SIGMOID
    return P


def fused_render(renderer, npc, decoders, rays_d, rays_o, stage, gt_depth, npc_geo_feats, npc_col_feats,
                 is_tracker, cloud_pos, dynamic_r_query, exposure_feat, far_group=None, n_surface=None,
                 force_save=False, return_ctx=False, feat_subset=None, z_zero_depth=None, save_light=False,
                 radius_override=None):
    """The one place that marshals a render call into lsr_render_fwd / lsr_render_bwd."""
    _lib.require_cuda(rays_o, 'rays_o')
    dev = rays_o.device
    if stage not in _lib.LSR_STAGE:
        raise NotImplementedError(f"stage '{stage}' is not on the fused path (only 'geometry' and 'color')")
    if not hasattr(decoders, 'blob'):
        raise TypeError('decoders must be loopy_slam_b200.decoder.NICER (flat weight blob); '
                        'build it with loopy_slam_b200.config.get_model(cfg)')
    prm = _make_params(renderer, decoders, stage, float(renderer.sigmoid_coefficient))
    if n_surface is not None:
        prm.n_surface = n_surface
    if save_light:
        prm.flags |= _lib.FLAG_SAVE_LIGHT
    radius = npc.get_radius_query() if npc is not None and hasattr(npc, 'get_radius_query') else renderer.radius_query
    if radius_override is not None:        # NICER.forward stage 'mesh': find_neighbors_faiss(step='mesh') -> radius_mesh
        radius = radius_override
    prm.radius_query = float(radius)
    rc = _RenderCtx()
    rc.prm = prm
    rc.renderer = renderer
    rc.stage = _lib.LSR_STAGE[stage]
    rc.is_tracker = bool(is_tracker)
    rc.r_query = None
    cell = float(radius)
    if renderer.use_dynamic_radius:
        if dynamic_r_query is None:
            raise ValueError('use_dynamic_radius is set but dynamic_r_query is None')
        rc.r_query = dynamic_r_query.detach().to(device=dev, dtype=torch.float64).reshape(-1).contiguous()
        cell = renderer.max_query_radius
    rc.grid = _grid_for(npc, cloud_pos, cell, renderer._grid_cache)
    rc.blob = decoders.blob
    rc.flat, rc.wstruct = rc.blob.ensure(dev)
    affine = None
    if decoders.cfg_flags['encode_exposure'] and rc.stage == 1:
        if exposure_feat is not None:      # decoder.py:535-540 (tracker): affine applied per sample
            affine = decoders.color_decoder.mlp_exposure(exposure_feat).to(torch.float32).contiguous()
            prm.rgb_mode = _lib.RGB_AFFINE_SIGMOID
        else:                              # decoder.py:541-542 (mapper): caller applies it after compositing
            prm.rgb_mode = _lib.RGB_RAW
    rays_o, rays_d = _f32c(rays_o), _f32c(rays_d)
    R = rays_o.shape[0]
    if gt_depth is None:       # Renderer.py:107-113: no sensor depth -> z in [near_end, 10]
        gt = torch.zeros(R, dtype=torch.float32, device=dev)
        far = torch.full((1,), 10.0, dtype=torch.float32, device=dev)
        fgroup = max(R, 1)
    else:
        gt = _f32c(gt_depth.detach().reshape(-1))
        fgroup = int(far_group) if far_group else max(R, 1)
        far = torch.empty((R + fgroup - 1) // fgroup if R else 1, dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            check(lib().lsr_far_bound(ptr(gt), R, fgroup, ptr(far), stream_ptr(dev)), 'lsr_far_bound')
    geo = _f32c(npc_geo_feats)
    col = _f32c(npc_col_feats) if npc_col_feats is not None else None
    params = rc.blob.tensors()
    rc.remap = geo_leaf = col_leaf = None
    if feat_subset is not None:
        subset, geo_leaf, col_leaf = feat_subset
        if npc_geo_feats.requires_grad or (npc_col_feats is not None and npc_col_feats.requires_grad):
            raise ValueError('with feat_subset the feature tables are constants: gradients go to the leaf blocks')
        if subset.n_points != geo.shape[0] or subset.remap.device != dev:
            raise ValueError('feat_subset was built for a different table size / device')
        if geo_leaf.shape != (len(subset), 32) or (col is not None and rc.stage == 1 and
                                                   (col_leaf is None or col_leaf.shape != (len(subset), 32))):
            raise ValueError('feat_subset leaf blocks must be (len(indices), 32)')
        rc.remap = subset.remap
        geo_leaf = _f32c(geo_leaf)
        col_leaf = _f32c(col_leaf) if col_leaf is not None else None
    rc.z_zero = None
    if z_zero_depth is not None:   # rendering.sample_near_pcl: (R, S) sample depths, rows of zero-depth rays are used
        rc.z_zero = _f32c(z_zero_depth.detach())
        if rc.z_zero.shape != (R, prm.n_surface):
            raise ValueError('z_zero_depth must be (n_rays, N_surface)')
        prm.flags |= _lib.FLAG_SAMPLE_NEAR_PCL
    rc.R = R
    rc.device = dev
    rc.far_group = fgroup
    rc.force_save = bool(force_save)
    rc.grad_enabled = torch.is_grad_enabled()
    rc.timing = getattr(renderer, '_timing', None)
    out = _RenderFn.apply(rc, rays_o, rays_d, gt, geo, col, affine, far, geo_leaf, col_leaf, *params)
    if return_ctx:
        return out, rc
    return out


def _saved_views(rc):
    """Python mirror of the head of lsr::saved_layout (csrc/lsr_render.cuh): float views of the planes a light save
    (LSR_FLAG_SAVE_LIGHT) holds -- k-NN results, occupancy logits, per-sample colours."""
    S = rc.prm.n_surface
    P = rc.R * S
    Pp = (P + 127) // 128 * 128 + 128
    f = rc.saved.view(torch.float32)
    o = 0
    v = {}

    def take(name, n, shape):
        nonlocal o
        v[name] = f[o:o + n].view(shape)
        o += n
    take('idx', Pp * 8, (Pp, 8)); take('w', Pp * 8, (Pp, 8)); take('D', Pp * 8, (Pp, 8)); take('misc', Pp * 4, (Pp, 4))
    take('occ', Pp, (Pp,))
    if rc.stage == 1:
        take('rgbs', Pp * 4, (Pp, 4)); take('outraw', Pp * 4, (Pp, 4))
    return v, P


def decode_points(decoders, p, npc, stage, npc_geo_feats, npc_col_feats, pts_num, cloud_pos, dynamic_r_query,
                  exposure_feat, renderer=None):
    """NICER.forward semantics (decoder.py:573-626), forward only: runs the fused kernel with one
    sample per "ray" (origin = the point, direction = 0, so the sample point is the origin exactly)
    and reads the per-sample occupancy / colour out of a LIGHT save (148 B per point), in chunks of
    renderer.points_batch_size points like Renderer.eval_points (Renderer.py:40-60).
    -> raw (P,4) [r,g,b,occ-logit], ray_mask (P/pts_num,) bool or None, point_mask (P,) bool.
    stage 'mesh' (:611-620) = 'color' with the mesher's radius (radius_mesh, neural_point.py:1694-1696) and no ray mask
    (:257-262); stage 'color_only' (:621-626) returns the (P,3) colour alone."""
    mesh, color_only = stage == 'mesh', stage == 'color_only'
    radius_override = None
    if mesh:
        radius_override = getattr(npc, 'radius_mesh', None)
        if radius_override is None:
            raise ValueError("stage 'mesh' needs npc.radius_mesh (pointcloud.radius_mesh)")
    if mesh or color_only:
        stage = 'color'
    if renderer is None:
        renderer = getattr(decoders, '_lsr_renderer', None)
        if renderer is None:
            raise RuntimeError('NICER.forward needs a Renderer: call Renderer.eval_points(...)')
    pts = _f32c(p.detach().reshape(-1, 3))
    raws, hass = [], []
    chunk = max(int(getattr(renderer, 'points_batch_size', 500000)), 1)
    with torch.no_grad():
        for b in range(0, max(pts.shape[0], 1), chunk):
            pb = pts[b:b + chunk]
            n = pb.shape[0]
            ones = torch.ones(n, dtype=torch.float32, device=pts.device)
            dyn = dynamic_r_query
            if dyn is not None and dyn.numel() == pts.shape[0]:
                dyn = dyn.reshape(-1)[b:b + chunk]
            _, rc = fused_render(renderer, npc, decoders, torch.zeros_like(pb), pb, stage, ones, npc_geo_feats,
                                 npc_col_feats, False, cloud_pos, dyn, exposure_feat, n_surface=1, force_save=True,
                                 return_ctx=True, save_light=True, radius_override=radius_override)
            v, m = _saved_views(rc)
            occ = v['occ'][:m].clone()
            hass.append(v['misc'][:m, 1] > 0.5)
            rgb = v['rgbs'][:m, :3].clone() if rc.stage == 1 else torch.zeros(m, 3, device=pts.device)
            raws.append(torch.cat([rgb, occ[:, None]], -1))
    raw, has = torch.cat(raws), torch.cat(hass)
    if color_only:
        return raw[:, :3]
    ray_mask = None
    if pts_num and not mesh:
        ray_mask = ~(has.view(-1, pts_num).sum(1) < int(decoders.cfg_flags['N_surface'] / 2 + 1))
    return raw, ray_mask, has


class Renderer(object):
    """Mirror of the reference Renderer (Renderer.py:6-22) -- same cfg keys, same attributes."""

    def __init__(self, cfg, args, slam, points_batch_size=500000, ray_batch_size=3000):
        self.ray_batch_size = ray_batch_size
        self.points_batch_size = points_batch_size
        self.N_surface = cfg['rendering']['N_surface']
        self.near_end_surface = cfg['rendering']['near_end_surface']
        self.far_end_surface = cfg['rendering']['far_end_surface']
        self.sample_near_pcl = cfg['rendering']['sample_near_pcl']
        self.skip_zero_depth_pixel = cfg['rendering']['skip_zero_depth_pixel']
        self.near_end = cfg['rendering']['near_end']
        self.use_dynamic_radius = cfg['use_dynamic_radius']
        self.crop_edge = 0 if cfg['cam']['crop_edge'] is None else cfg['cam']['crop_edge']
        self.H, self.W, self.fx, self.fy, self.cx, self.cy = slam.H, slam.W, slam.fx, slam.fy, slam.cx, slam.cy
        self.radius_query = cfg['pointcloud']['radius_query']
        pc = cfg['pointcloud']
        # largest radius the dynamic-radius map can produce (Tracker.py:243-258): r_add_max * ratio
        self.max_query_radius = float(pc.get('radius_add_max', 0.08)) * float(pc.get('radius_query_ratio', 2))
        self.sigmoid_coefficient = cfg['rendering'].get('sigmoid_coef_mapper', 0.1)   # callers overwrite it
        self._grid_cache = _GridCache()

    # -- Renderer.py:24-69
    def eval_points(self, p, decoders, npc, stage='color', device=None, npc_geo_feats=None, npc_col_feats=None,
                    is_tracker=False, cloud_pos=None, pts_views_d=None, ray_pts_num=None, dynamic_r_query=None,
                    exposure_feat=None):
        assert torch.is_tensor(p)
        raw, ray_mask, point_mask = decode_points(decoders, p, npc, stage, npc_geo_feats, npc_col_feats,
                                                  ray_pts_num, cloud_pos, dynamic_r_query, exposure_feat,
                                                  renderer=self)
        return raw, ray_mask, point_mask

    # -- Renderer.py:71-201
    def render_batch_ray(self, npc, decoders, rays_d, rays_o, device, stage, gt_depth=None, npc_geo_feats=None,
                         npc_col_feats=None, is_tracker=False, cloud_pos=None, dynamic_r_query=None,
                         exposure_feat=None, feat_subset=None):
        """-> depth (R,), uncertainty (R,), color (R,3), valid_ray_mask (R,) bool; differentiable w.r.t.
        rays, feature tables, decoder parameters and exposure_feat.  feat_subset (extension, optional):
        (FeatureSubset, geo_leaf, col_leaf) -- see FeatureSubset."""
        if gt_depth is not None and torch.numel(gt_depth) == 0:
            warnings.warn('tensor gt_depth is empty, info:')      # Renderer.py:122-128
            gt_depth = None
        if self.sample_near_pcl and hasattr(npc, 'sample_near_pcl'):
            return self._render_with_near_pcl(npc, decoders, rays_d, rays_o, stage, gt_depth, npc_geo_feats,
                                              npc_col_feats, is_tracker, cloud_pos, dynamic_r_query, exposure_feat,
                                              feat_subset)
        depth, var, rgb, valid = fused_render(self, npc, decoders, rays_d, rays_o, stage, gt_depth, npc_geo_feats,
                                              npc_col_feats, is_tracker, cloud_pos, dynamic_r_query, exposure_feat,
                                              feat_subset=feat_subset)
        return depth, var, rgb, valid.bool()

    def _render_with_near_pcl(self, npc, decoders, rays_d, rays_o, stage, gt_depth, npc_geo_feats, npc_col_feats,
                              is_tracker, cloud_pos, dynamic_r_query, exposure_feat, feat_subset):
        """Renderer.py:150-158,191-198 with rendering.sample_near_pcl: rays without sensor depth take their samples
        from npc.sample_near_pcl (between the first two of 25 coarse steps that have a neighbour), keep their
        rendered depth, and are invalid when fewer than two such steps exist."""
        z_zero = None
        not_near_rays = None
        if gt_depth is None:                 # Renderer.py:107-113: all rays zero-depth, far = 10
            z0, not_near = npc.sample_near_pcl(rays_o.clone().detach(), rays_d.clone().detach(), self.near_end, 10.0,
                                               self.N_surface)
            z_zero = z0.float()
            not_near_rays = torch.nonzero(not_near, as_tuple=True)[0]
            zero = None
        else:
            g = gt_depth.detach().reshape(-1)
            zero = ~(g > 0)
        if zero is not None and bool(zero.any()):
            R = g.shape[0]
            g32 = _f32c(g)
            far = torch.empty(1, dtype=torch.float32, device=g32.device)
            with _lib.on_device(g32.device):
                check(lib().lsr_far_bound(ptr(g32), R, max(R, 1), ptr(far), stream_ptr(g32.device)), 'lsr_far_bound')
            z0, not_near = npc.sample_near_pcl(rays_o[zero].clone().detach(), rays_d[zero].clone().detach(),
                                               self.near_end, float(far.item()), self.N_surface)   # Renderer.py:151-153
            z_zero = torch.zeros(R, self.N_surface, dtype=torch.float32, device=g32.device)
            z_zero[zero] = z0.to(device=g32.device, dtype=torch.float32)
            not_near_rays = torch.nonzero(zero, as_tuple=True)[0][not_near.to(g32.device)]
        depth, var, rgb, valid = fused_render(self, npc, decoders, rays_d, rays_o, stage, gt_depth, npc_geo_feats,
                                              npc_col_feats, is_tracker, cloud_pos, dynamic_r_query, exposure_feat,
                                              feat_subset=feat_subset, z_zero_depth=z_zero)
        valid = valid.bool()
        if not_near_rays is not None and not_near_rays.numel():
            valid = valid.clone()
            valid[not_near_rays] = False                                                        # Renderer.py:154-157,194
        return depth, var, rgb, valid

    def _near_pcl_depths_img(self, npc, rays_o, rays_d, gt):
        """Sample depths of the zero-depth pixels of a full image (Renderer.py:150-158 applied per ray_batch_size tile, as
        the reference's tile loop :243-266 does: each tile uses ITS far bound).  Only tiles that hold a zero-depth pixel are
        visited; -> (H*W, N_surface) or None when every pixel has sensor depth."""
        R = rays_o.shape[0]
        if gt is None:                       # Renderer.py:107-113: every ray is zero-depth, far = 10 for every tile
            z0, _ = npc.sample_near_pcl(rays_o, rays_d, self.near_end, 10.0, self.N_surface)
            return z0.float()
        g32 = _f32c(gt.detach())
        zero = ~(g32 > 0)
        if not bool(zero.any()):
            return None
        G = self.ray_batch_size
        far = torch.empty((R + G - 1) // G, dtype=torch.float32, device=g32.device)
        with _lib.on_device(g32.device):
            check(lib().lsr_far_bound(ptr(g32), R, G, ptr(far), stream_ptr(g32.device)), 'lsr_far_bound')
        far_h = far.cpu()
        z_zero = torch.zeros(R, self.N_surface, dtype=torch.float32, device=g32.device)
        tiles = torch.unique(torch.nonzero(zero, as_tuple=True)[0] // G).cpu().tolist()
        for t in tiles:
            sl = slice(t * G, min((t + 1) * G, R))
            zt = zero[sl]
            z0, _ = npc.sample_near_pcl(rays_o[sl][zt], rays_d[sl][zt], self.near_end, float(far_h[t]), self.N_surface)
            z_zero[sl][zt] = z0.float()
        return z_zero

    # -- Renderer.py:203-276
    def render_img(self, npc, decoders, c2w, device, stage, gt_depth=None, npc_geo_feats=None, npc_col_feats=None,
                   dynamic_r_query=None, cloud_pos=None, exposure_feat=None):
        """Full-image forward render in ONE fused launch (the reference loops over 3000-ray tiles).
        -> depth (H,W) f64, uncertainty (H,W) f64, color (H,W,3) f32."""
        from .common import get_rays
        with torch.no_grad():
            H, W = self.H, self.W
            rays_o, rays_d = get_rays(H, W, self.fx, self.fy, self.cx, self.cy, c2w, device)
            rays_o = rays_o.reshape(-1, 3)
            rays_d = rays_d.reshape(-1, 3)
            dyn = dynamic_r_query.reshape(-1) if (self.use_dynamic_radius and dynamic_r_query is not None) else None
            gt = gt_depth.reshape(-1) if gt_depth is not None else None
            z_zero = None
            if self.sample_near_pcl and hasattr(npc, 'sample_near_pcl'):
                z_zero = self._near_pcl_depths_img(npc, rays_o, rays_d, gt)
            depth, var, rgb, _ = fused_render(self, npc, decoders, rays_d, rays_o, stage, gt, npc_geo_feats,
                                              npc_col_feats, False, cloud_pos, dyn, exposure_feat,
                                              far_group=self.ray_batch_size, z_zero_depth=z_zero)
            return depth.double().reshape(H, W), var.double().reshape(H, W), rgb.reshape(H, W, 3)
