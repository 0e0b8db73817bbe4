#!/usr/bin/env python
"""Headline benchmark: rendered rays/s (forward + backward) of Loopy-SLAM's per-iteration render hot path on a
synthetic Replica-room0-shaped stream (BASELINE.json metric).

One "step" = ONE MAPPING ITERATION exactly as the reference's loop body issues it
(/root/reference/src/Mapper.py:576-735, optimiser step excluded):

    npc_geo_feats[indices] = geo_pcl_grad ; npc_col_feats[indices] = color_pcl_grad      Mapper.py:581-582
    for each of the 12 window frames: get_samples(0, H, 0, W, pixels // n_frames, ..., c2w, depth, colour,
                                                  depth_filter=True, return_index=True)   Mapper.py:652-655
    cat, inside_mask (<= min(10 median, 1.2 max) of gt depth)                             Mapper.py:674-681
    Renderer.render_batch_ray(..., stage, gt_depth, npc_geo_feats, npc_col_feats, cloud_pos)   Mapper.py:682-688
    masked L1 depth (+ w * L1 colour) as inline torch ops                                 Mapper.py:689-720
    loss.backward()                                                                       Mapper.py:722
    tables .detach()                                                                      Mapper.py:727-735

  value   : device-timed (CUDA events around the step), keyframe images, cloud and weights resident in HBM.
  e2e     : the same public calls, wall clock: every step uploads the window's camera poses from PINNED host memory
            and reads the loss back; the current frame's RGB-D image is uploaded once inside the timed region (the
            reference uploads a frame once per mapped frame = once per >= 300 iterations; the keyframe images stay on
            the device, src/Mapper.py:637-638).
  extra   : the same iteration with the library's caller-side extensions (row_remap instead of the index_put,
            fused loss kernel), the geometry stage, a tracking iteration, the TUM / ScanNet shaped iterations
            (dynamic radius, exposure) and the N sweep -- each labelled.
  --impl reference : the reference's own CPU path.  The reference is pure Python / PyTorch + faiss-gpu and cannot
            travel to or run on this box; timed: the oracle restatement of the SAME iteration (oracle/
            sampling.py + oracle/render.py + exact C grid k-NN) on all host threads.
Multi-GPU (torchrun, one rank per GPU): the rays of an iteration shard across ranks -- every rank samples its own
pixels of the window frames -- then ONE NCCL all-reduce over [d_geo | d_col | decoder grads] per step.
--scaling weak : pixels per rank fixed (default, = the driver's scaling run);  strong: the reference's fixed
iteration (mapping.pixels rays in total) split over the ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

T_START = time.time()

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS, SEED = 200000, 1219
S, K, C = 5, 8, 32
MFLOP_RAY_FWD = {('color', True): 1.99, ('color', False): 1.13, ('geometry', True): 0.157, ('geometry', False): 0.157}
METRIC = 'rendered rays/sec (fwd+bwd)'


def bytes_ray(stage, fwd_only=False):
    T = 2 if stage == 'color' else 1
    return 53 + S * (K * 12 + T * K * C * 4 * (1 if fwd_only else 3))


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)', d
    return 6650.0, 'fallback (B200_PROFILING.md)', {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (one streaming `nvidia-smi -lms` subprocess)."""

    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0, period_ms=20):
        self.rows, self.keep, self.proc = [], False, None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', str(period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _run(self):
        for line in self.proc.stdout:
            if self.keep:
                self.rows.append([x.strip() for x in line.split(',')])

    def start(self):
        self.keep = True

    def stop(self):
        self.keep = False
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        sm = sorted(int(r[0]) for r in rows)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith('active') for r in rows)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': int(rows[0][1]) if rows[0][1].isdigit() else None,
                'reasons': reasons, 'samples': len(rows)}


# ------------------------------------------------------------------------------ workload (shared by both arms)
def make_cfg(dataset):
    import loopy_slam_b200 as L
    return L.default_cfg(dataset)


def workload_config(dataset, stage, n_points, world, pixels_total, n_frames, H, W, scaling):
    """The `config` object of the JSON line: IDENTICAL keys and values in both arms."""
    flags = {'replica': 'Replica decoder flags (rel-pos neighbour MLP on, fixed radius)',
             'tum': 'TUM flags (dynamic radius, rel-pos off)',
             'scannet': 'ScanNet flags (dynamic radius, exposure on, band 0.96-1.04)'}[dataset]
    return {'workload': f"{dataset} default config, one mapping iteration (src/Mapper.py:576-735 without optimizer.step), "
                        f"stage '{stage}', {flags}",
            'pixels_per_step': pixels_total, 'frames_per_step': n_frames, 'n_surface': S, 'n_points': int(n_points),
            'image': f'{H}x{W}', 'scaling': scaling, 'n_gpus': world,
            'l2': 'flushed between timed steps (256 MiB memset outside the event pairs)'}


def build_scene(dataset, n_points, cache=True):
    """Synthetic stream + cloud on the host.  -> dict(cloud, geo, col, frames=[(color, depth, c2w)], room)."""
    from loopy_slam_b200.stream import SyntheticRoom, build_point_cloud
    cfg = make_cfg(dataset)
    cam = cfg['cam']
    ce = cam['crop_edge'] or 0
    room = SyntheticRoom(H=cam['H'] - 2 * ce, W=cam['W'] - 2 * ce, fx=cam['fx'], fy=cam['fy'], cx=cam['cx'] - ce,
                         cy=cam['cy'] - ce, seed=SEED)
    path = f'/tmp/lsr_bench_cloud_{dataset}_{n_points}.pt'
    if cache and os.path.exists(path):
        cloud, geo, col = torch.load(path)
    elif n_points > 300000:
        # The insertion rule saturates near 2.3e5 points in this room (locations >= radius_add apart on ~100 m^2 of wall), so
        # the large-N sweep point is the 2e5 cloud replicated with a 1 cm jitter: same geometry, proportionally denser cells.
        base = build_scene(dataset, N_POINTS, cache)
        reps = (n_points + base['cloud'].shape[0] - 1) // base['cloud'].shape[0]
        gen = torch.Generator().manual_seed(SEED + 7)
        cloud = torch.cat([base['cloud'] + (0.01 * torch.randn(base['cloud'].shape, generator=gen) if r else 0.0)
                           for r in range(reps)])[:n_points].contiguous()
        geo = torch.zeros(cloud.shape[0], 32).normal_(0, 0.1, generator=gen)
        col = torch.zeros(cloud.shape[0], 32).normal_(0, 0.1, generator=gen)
    else:
        cloud, geo, col = build_point_cloud(room, n_points, frame_stride=40, pixels_per_frame=20000, seed=SEED, max_frames=400)
    if not (cache and os.path.exists(path)):
        if cache:
            try:
                torch.save((cloud, geo, col), path + f'.{os.getpid()}')
                os.replace(path + f'.{os.getpid()}', path)
            except Exception:
                pass
    n_frames = cfg['mapping']['mapping_window_size']
    frames = [room.frame(40 * f) for f in range(n_frames)]
    return dict(cloud=cloud, geo=geo, col=col, frames=frames, room=room, cfg=cfg)


def mapper_loss_reference(depth, color, valid_ray_mask, batch_gt_depth, batch_gt_color, stage, w_color_loss):
    """src/Mapper.py:689-693 + 716-720, verbatim structure (boolean-mask indexing and all)."""
    depth_mask = (batch_gt_depth > 0) & valid_ray_mask
    depth_mask = depth_mask & (~torch.isnan(depth))
    geo_loss = torch.abs(batch_gt_depth[depth_mask] - depth[depth_mask]).sum()
    loss = geo_loss.clone()
    if stage == 'color':
        color_loss = torch.abs(batch_gt_color[depth_mask] - color[depth_mask]).sum()
        loss = loss + w_color_loss * color_loss
    return loss


class MapperIteration:
    """One mapping iteration on the device through loopy_slam_b200's mirror of the reference surface."""

    def __init__(self, dataset, n_points, dev, rank=0, world=1, pixels_total=None, stage='color'):
        import loopy_slam_b200 as L
        from loopy_slam_b200.frustum import get_mask_from_c2w
        from loopy_slam_b200.radius_map import dynamic_radius_maps
        self.L = L
        self.dev = dev
        self.stage = stage
        sc = build_scene(dataset, n_points)
        self.sc = sc
        cfg = sc['cfg']
        self.cfg = cfg
        room = sc['room']
        self.room = room
        self.n_frames = len(sc['frames'])
        self.pixels_total = pixels_total or cfg['mapping']['pixels']
        self.pix_per_image = self.pixels_total // self.n_frames                 # Mapper.py:466
        torch.manual_seed(SEED)
        self.model = L.get_model(cfg).to(dev)
        torch.manual_seed(SEED + 1000 * rank)       # every rank draws its own pixels (torch.randint in get_samples)

        class Slam:
            H, W, fx, fy, cx, cy = room.H, room.W, room.fx, room.fy, room.cx, room.cy
        self.rend = L.Renderer(cfg, None, Slam)
        self.rend.sigmoid_coefficient = cfg['rendering']['sigmoid_coef_mapper']
        rq = cfg['pointcloud']['radius_query']

        class NPC:
            def get_radius_query(self):
                return rq
        self.npc = NPC()
        self.cloud = sc['cloud'].to(dev)
        self.npc_geo, self.npc_col = sc['geo'].to(dev), sc['col'].to(dev)
        self.colors = [f[0].to(dev) for f in sc['frames']]
        self.depths = [f[1].to(dev) for f in sc['frames']]
        self.c2ws = [f[2].to(dev) for f in sc['frames']]
        self.dyn = cfg['use_dynamic_radius']
        self.r_maps = [dynamic_radius_maps(c, cfg)[1] for c in self.colors] if self.dyn else None
        self.exposure = cfg['model']['encode_exposure']
        self.exposure_feats = [torch.zeros(cfg['model']['exposure_dim'], device=dev).normal_(0, 0.01).requires_grad_(True)
                               for _ in range(self.n_frames)] if self.exposure else None
        self.indices = get_mask_from_c2w(self.cloud, sc['frames'][-1][2], self.depths[-1], room.H, room.W, room.fx, room.fy,
                                         room.cx, room.cy, edge=-4)                      # Mapper.py:498-500
        self.geo_leaf = self.npc_geo[self.indices].clone().requires_grad_(True)          # Mapper.py:502-505
        self.col_leaf = self.npc_col[self.indices].clone().requires_grad_(True)
        # mapping.fix_geo_decoder: colour decoder + the Fourier matrices of the geometry decoder train (Mapper.py:524-541)
        for p in self.model.geo_decoder.parameters():
            p.requires_grad_(False)
        self.model.geo_decoder.embedder._B.requires_grad_(True)
        self.train_params = [p for p in self.model.parameters() if p.requires_grad] + [self.geo_leaf, self.col_leaf] + \
                            (self.exposure_feats or [])
        self.subset = L.FeatureSubset(self.indices, self.npc_geo.shape[0])
        self.w_color = cfg['mapping']['w_color_loss']
        self.rays_last = 0

    def sample(self, c2ws=None):
        """Mapper.py:624-681: the window's rays."""
        L, room, dev = self.L, self.room, self.dev
        O, D, G, Cc, Rq, counts = [], [], [], [], [], []
        for f in range(self.n_frames):
            c2w = self.c2ws[f] if c2ws is None else c2ws[f]
            o, d, g, c, i, j = L.get_samples(0, room.H, 0, room.W, self.pix_per_image, room.H, room.W, room.fx, room.fy,
                                             room.cx, room.cy, c2w, self.depths[f], self.colors[f], dev,
                                             depth_filter=True, return_index=True)
            O.append(o.float()); D.append(d.float()); G.append(g.float()); Cc.append(c.float())
            counts.append(o.shape[0])
            if self.dyn:
                Rq.append(self.r_maps[f][j, i])
        o, d, g, c = torch.cat(O), torch.cat(D), torch.cat(G), torch.cat(Cc)
        rq = torch.cat(Rq) if self.dyn else None
        with torch.no_grad():
            inside = g <= torch.minimum(10 * g.median(), 1.2 * torch.max(g))
        keep = torch.nonzero(inside, as_tuple=True)[0]
        frame_id = None
        if self.exposure:
            frame_id = torch.repeat_interleave(torch.arange(self.n_frames, device=dev),
                                               torch.tensor(counts, device=dev))[keep]
        o, d, g, c = o[keep], d[keep], g[keep], c[keep]
        if rq is not None:
            rq = rq[keep]
        return o, d, g, c, rq, frame_id

    def apply_exposure(self, color, frame_id):
        """Mapper.py:697-715: per-frame affine colour transform after compositing, then the sigmoid."""
        counts = torch.bincount(frame_id, minlength=self.n_frames).tolist()
        out, start = [], 0
        for f, n in enumerate(counts):
            aff = self.model.color_decoder.mlp_exposure(self.exposure_feats[f])
            rot, trans = aff[:9].reshape(3, 3), aff[-3:]
            out.append(torch.matmul(color[start:start + n], rot) + trans)
            start += n
        return torch.sigmoid(torch.cat(out))

    def step(self, literal=True, c2ws=None, stage=None, timing=None):
        L, dev = self.L, self.dev
        stage = stage or self.stage
        for p in self.train_params:
            p.grad = None
        self.rend._timing = timing
        o, d, g, c, rq, frame_id = self.sample(c2ws)
        self.rays_last = o.shape[0]
        if literal:      # Mapper.py:581-582: the leaf blocks are written into the full tables every iteration
            self.npc_geo[self.indices] = self.geo_leaf
            self.npc_col[self.indices] = self.col_leaf
            depth, var, color, valid = self.rend.render_batch_ray(self.npc, self.model, d, o, dev, stage, gt_depth=g,
                                                                  npc_geo_feats=self.npc_geo, npc_col_feats=self.npc_col,
                                                                  is_tracker=False, cloud_pos=self.cloud, dynamic_r_query=rq,
                                                                  exposure_feat=None)
        else:            # library extension: same rows through row_remap, leaf-sized gradients, no table rewrite
            depth, var, color, valid = self.rend.render_batch_ray(self.npc, self.model, d, o, dev, stage, gt_depth=g,
                                                                  npc_geo_feats=self.npc_geo, npc_col_feats=self.npc_col,
                                                                  is_tracker=False, cloud_pos=self.cloud, dynamic_r_query=rq,
                                                                  exposure_feat=None,
                                                                  feat_subset=(self.subset, self.geo_leaf, self.col_leaf))
        if self.exposure and stage == 'color':
            color = self.apply_exposure(color, frame_id)
        if literal:
            loss = mapper_loss_reference(depth, color, valid, g, c, stage, self.w_color)
        else:
            loss = L.mapper_loss(depth, color, valid, g, c, stage, self.w_color)[0]
        loss.backward()
        if literal:      # Mapper.py:727-735
            self.npc_geo, self.npc_col = self.npc_geo.detach(), self.npc_col.detach()
        self.rend._timing = None
        return loss


def time_steps(fn, n, flush, dev, warm=3):
    """mean device ms of fn(k) over n calls (CUDA events, L2 flushed in between); fn returns the rays it processed."""
    for k in range(warm):
        fn(k)
    torch.cuda.synchronize(dev)
    evs, rays = [], 0
    for k in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rays += fn(k)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize(dev)
    ms = sum(a.elapsed_time(b) for a, b in evs) / n
    return ms, rays / n


def run_lsr(args, rank, world, local):
    import loopy_slam_b200 as L
    from loopy_slam_b200 import parallel, _lib
    import __graft_entry__ as entry
    if not os.path.exists(_lib.LIB_PATH):
        entry.build()
    dev = torch.device(f'cuda:{local}')
    torch.cuda.set_device(dev)
    stage = args.stage
    cfg0 = make_cfg(args.config)
    pixels_total = cfg0['mapping']['pixels']
    if args.scaling == 'strong' and world > 1:
        pixels_rank = pixels_total // world
    else:
        pixels_rank = pixels_total
    it = MapperIteration(args.config, args.n_points, dev, rank, world, pixels_rank, stage)
    reducer = parallel.GradAllReducer(it.train_params)
    lib = _lib.lib()
    timing = {'fwd': [], 'bwd': []}
    rays_seen = [0]

    def step(literal=True, c2ws=None, timed=False):
        loss = it.step(literal=literal, c2ws=c2ws, timing=timing if timed else None)
        if world > 1:
            reducer.allreduce_(getattr(it.rend, 'last_grad_buffer', None))
        rays_seen[0] += it.rays_last
        return loss

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local) if rank == 0 else None
    for w in range(args.warmup):
        step()
    barrier()
    if sampler:
        sampler.start()
    lib.lsr_launch_count(1)
    rays_seen[0] = 0
    evs = []
    for k in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(timed=True)
        e1.record()
        evs.append((e0, e1))
    barrier()
    n_launch = int(lib.lsr_launch_count(1))
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    ms_step = parallel.max_over_ranks(ms_total / args.steps, dev)
    rays_step = rays_seen[0] / args.steps                                        # this rank's mean rays per step
    rays_all = parallel.max_over_ranks(rays_step, dev) * world if world > 1 else rays_step
    kt = {k: [a.elapsed_time(b) for a, b in v] for k, v in timing.items()}

    # ---- e2e: poses from pinned host memory every step, the current frame's RGB-D once, loss read back; wall clock
    poses_host = torch.stack([f[2] for f in it.sc['frames']]).pin_memory()       # (n_frames, 4, 4)
    cur_color_h = it.sc['frames'][-1][0].pin_memory()
    cur_depth_h = it.sc['frames'][-1][1].pin_memory()

    def e2e_step(upload_frame=False):
        if upload_frame:
            it.colors[-1] = cur_color_h.to(dev, non_blocking=True)
            it.depths[-1] = cur_depth_h.to(dev, non_blocking=True)
        poses = poses_host.to(dev, non_blocking=True)
        loss = step(c2ws=[poses[f] for f in range(it.n_frames)])
        return loss.item()                                                       # D2H of the step's result
    for w in range(2):
        e2e_step()
    barrier()
    rays_seen[0] = 0
    t0 = time.perf_counter()
    for k in range(args.steps):
        loss_host = e2e_step(upload_frame=(k == 0))
    barrier()
    e2e_ms = parallel.max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps, dev)
    e2e_rays = rays_seen[0] / args.steps
    e2e_rays_all = parallel.max_over_ranks(e2e_rays, dev) * world if world > 1 else e2e_rays
    clocks = sampler.stop() if sampler else None
    frame_bytes = cur_color_h.numel() * 4 + cur_depth_h.numel() * 4
    h2d_bytes = poses_host.numel() * 4 + frame_bytes / args.steps

    extra = {}
    if world == 1 and not args.no_extra:
        n_extra = max(10, args.steps // 4)

        def line(ms, rays, **kw):
            return dict({'rays_per_s': rays / (ms * 1e-3), 'ms_per_step': ms, 'rays_per_step': rays}, **kw)
        # (1) the same iteration with the library's extensions (row_remap + fused loss kernel)
        ms, r = time_steps(lambda k: (it.step(literal=False), it.rays_last)[1], n_extra, flush, dev)
        extra['mapper_iteration_lsr_extensions'] = line(ms, r, what='FeatureSubset/row_remap instead of the per-iteration index_put, '
                                                        'lsr_mapper_loss instead of the inline torch loss (callers must opt in)')
        # (1b) the render hot path alone: rays of eight iterations pre-sampled and resident on the device, row_remap + fused
        # loss, no host synchronisation inside the step -- what the kernels sustain when the caller does not serialise them
        with torch.no_grad():
            pre = [it.sample() for _ in range(8)]

        def hot(k):
            o, d, g, c, rq, _ = pre[k % 8]
            for p in it.train_params:
                p.grad = None
            depth, var, color, valid = it.rend.render_batch_ray(it.npc, it.model, d, o, dev, stage, gt_depth=g,
                                                                npc_geo_feats=it.npc_geo, npc_col_feats=it.npc_col,
                                                                is_tracker=False, cloud_pos=it.cloud, dynamic_r_query=rq,
                                                                feat_subset=(it.subset, it.geo_leaf, it.col_leaf))
            it.L.mapper_loss(depth, color, valid, g, c, stage, it.w_color)[0].backward()
            return o.shape[0]
        if not it.exposure:
            ms, r = time_steps(hot, n_extra, flush, dev)
            extra['render_hot_path_only'] = line(ms, r, what='render_batch_ray + fused loss + backward on pre-sampled device-resident '
                                                 'rays (row_remap), no host synchronisation in the step')
        # (2) geometry stage (40 % of the mapping iterations, Mapper.py:588-591)
        ms, r = time_steps(lambda k: (it.step(stage='geometry'), it.rays_last)[1], n_extra, flush, dev)
        extra['mapper_iteration_geometry_stage'] = line(ms, r)
        # (3) tracking iteration (src/Tracker.py:102-197): frozen decoders, gradient to the pose through the rays
        extra['tracker_iteration'] = tracker_line(it, cfg0, n_extra, flush, dev)
        # (4) full-image forward render (Renderer.render_img, src/utils/Renderer.py:203-276)
        with torch.no_grad():
            def img(k):
                it.rend.render_img(it.npc, it.model, it.c2ws[-1], dev, 'color', gt_depth=it.depths[-1], npc_geo_feats=it.npc_geo,
                                   npc_col_feats=it.npc_col, cloud_pos=it.cloud,
                                   dynamic_r_query=it.r_maps[-1] if it.dyn else None)
                return it.room.H * it.room.W
            ms, r = time_steps(img, 5, flush, dev, warm=1)
        extra['render_img_forward_only'] = line(ms, r)
        # (5) the other BASELINE configs and the N sweep, same literal iteration
        if not args.no_sweep:
            for ds, npts in (('tum', N_POINTS), ('scannet', N_POINTS), (args.config, 2000), (args.config, 1000000)):
                if ds == args.config and npts == args.n_points:
                    continue
                if time.time() - T_START > args.sweep_budget_s:      # keep the default run within minutes
                    extra[f'mapper_iteration_{ds}_N{npts}'] = {'skipped': f'wall-clock budget of {args.sweep_budget_s} s for side measurements spent'}
                    continue
                try:
                    it2 = MapperIteration(ds, npts, dev, 0, 1, None, 'color')
                    ms, r = time_steps(lambda k: (it2.step(), it2.rays_last)[1], n_extra, flush, dev)
                    extra[f'mapper_iteration_{ds}_N{npts}'] = line(ms, r, n_points=int(it2.cloud.shape[0]),
                                                                   config=workload_config(ds, 'color', it2.cloud.shape[0], 1,
                                                                                          it2.pixels_total, it2.n_frames,
                                                                                          it2.room.H, it2.room.W, 'weak'))
                    del it2
                    torch.cuda.empty_cache()
                except Exception as e:   # a side measurement must not take the headline down
                    extra[f'mapper_iteration_{ds}_N{npts}'] = {'error': repr(e)[:200]}

    if rank != 0:
        return
    peak, peak_src, peaks = load_peaks()
    mean = lambda xs: sum(xs) / max(len(xs), 1)
    t_f, t_b = mean(kt['fwd']), mean(kt['bwd'])
    relpos = bool(it.cfg['model']['encode_rel_pos_in_col'])
    R = rays_step
    dom_bwd = t_b >= t_f
    bytes_dom = R * ((bytes_ray(stage) - bytes_ray(stage, True)) if dom_bwd else bytes_ray(stage, True))
    t_dom = max(t_b, t_f)
    ach = bytes_dom / (t_dom * 1e-3) / 1e9 if t_dom > 0 else 0.0
    sm_mhz = (clocks or {}).get('sm_mhz') or 1965
    # tcgen05 kind::tf32 rate measured on this pool (tools/umma_rate_probe.cu, profiles/r02_probe_mma_rate.log):
    # 65.4 cycles per M128 N128 K8 MMA per SM  =>  2*128*128*8 / 65.4 FLOP per cycle per SM
    tf32_peak = 148 * (2 * 128 * 128 * 8 / 65.4) * sm_mhz * 1e6 / 1e12
    flops_step = R * MFLOP_RAY_FWD[(stage, relpos)] * 1e6 * 3
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'r02_ncu_traffic.json')
    if os.path.exists(tp) and args.config == 'replica' and stage == 'color':
        traffic = json.load(open(tp))
    out = {
        'metric': METRIC, 'value': rays_all / (ms_step * 1e-3), 'unit': 'rays/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.config, stage, it.cloud.shape[0], world, pixels_total, it.n_frames, it.room.H, it.room.W,
                                  args.scaling),
        'rays_per_step': rays_all, 'n_trainable_rows': int(it.indices.shape[0]),
        'parallelism': f'ray-shard dp{world} (replicated cloud + weights, every rank samples its own pixels, 1 NCCL all-reduce/step)',
        'e2e': {'value': e2e_rays_all / (e2e_ms * 1e-3), 'unit': 'rays/s', 'ms_per_step': e2e_ms,
                'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4,
                'note': f'{poses_host.numel() * 4} B of poses every step + the current RGB-D frame ({frame_bytes} B) once inside the '
                        f'timed region; keyframe images stay on the device as in the reference'},
        'gpu_launches': n_launch,
        'roofline': {'bound': 'hbm', 'kernel': 'lsr_render_bwd (trunk_bwd_umma + geo_bwd_umma + render_bwd and their prep / trig / finalize kernels, two streams)' if dom_bwd
                     else 'lsr_render_fwd (sample_knn + render_fwd kernels)',
                     'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                     'traffic': (traffic or {}).get('bwd_bytes' if dom_bwd else 'fwd_bytes'), 'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': bytes_dom,
                     'traffic_source': (traffic or {}).get('source'),
                     'note': 'HBM fraction as BASELINE.md section 3 defines it; at ~190 FLOP/B both passes sit on the tensor/issue '
                             'side of the roofline, see "tensor"'},
        'tensor': {'algorithmic_tflops': flops_step / ((t_f + t_b) * 1e-3) / 1e12 if t_f + t_b > 0 else 0,
                   'executed_tflops_3xtf32': 3 * flops_step / ((t_f + t_b) * 1e-3) / 1e12 if t_f + t_b > 0 else 0,
                   'peak_tf32_tcgen05_tflops_measured': tf32_peak, 'peak_bf16_tflops_measured': peaks.get('bf16_tflops'),
                   'frac_of_tf32_peak_executed': (3 * flops_step / ((t_f + t_b) * 1e-3) / 1e12) / tf32_peak if t_f + t_b > 0 else 0,
                   'note': 'peak = measured tcgen05 kind::tf32 issue rate (65.4 cycles per M128 N128 K8 MMA per SM, '
                           'profiles/r02_probe_mma_rate.log) at the sampled SM clock; 3 tensor-core passes per fp32 product '
                           '(hi*lo + lo*hi + hi*hi) keep the 1e-4 parity contract'},
        'kernels': {'lsr_render_fwd_ms': t_f, 'lsr_render_bwd_ms': t_b,
                    'fwd_bwd_GBps': R * bytes_ray(stage) / ((t_f + t_b) * 1e-3) / 1e9 if t_f + t_b > 0 else 0},
        'extra': extra, 'clocks': clocks, 'loss': loss_host,
    }
    if world == 1 and not args.no_cpu_baseline and args.config == 'replica':
        out['cpu_baseline'] = cpu_baseline(args, it)
        out['depth_l1_vs_reference_m'] = out['cpu_baseline']['parity']['depth_l1_vs_oracle_m']
    print(json.dumps(out))


def c2w_to_cam7(c2w):
    """4x4 camera-to-world -> [quaternion (w, x, y, z) | translation] (the tracker's pose parametrisation)."""
    R = c2w[:3, :3].double()
    t = R[0, 0] + R[1, 1] + R[2, 2]
    w = torch.sqrt(torch.clamp(1 + t, min=1e-12)) / 2
    x = (R[2, 1] - R[1, 2]) / (4 * w)
    y = (R[0, 2] - R[2, 0]) / (4 * w)
    z = (R[1, 0] - R[0, 1]) / (4 * w)
    return torch.cat([torch.stack([w, x, y, z]), c2w[:3, 3].double()]).float()


def tracker_line(it, cfg, n, flush, dev):
    """src/Tracker.py:102-197: camera tensor -> c2w -> get_samples inside the border -> render (is_tracker) -> tracker
    loss -> backward to the 7 pose parameters."""
    L = it.L
    room = it.room
    for p in it.model.parameters():
        p.requires_grad_(False)
    trk = cfg['tracking']
    Wedge, Hedge, pixels = trk['ignore_edge_W'], trk['ignore_edge_H'], trk['pixels']
    c2w = it.sc['frames'][-1][2]
    cam = c2w_to_cam7(c2w).to(dev).requires_grad_(True)
    it.rend.sigmoid_coefficient = cfg['rendering']['sigmoid_coef_tracker']
    rays = [0]

    def trk_step(k=0):
        cam.grad = None
        c = L.get_camera_from_tensor(cam)
        o, d, g, col = L.get_samples(Hedge, room.H - Hedge, Wedge, room.W - Wedge, pixels, room.H, room.W, room.fx, room.fy,
                                     room.cx, room.cy, c, it.depths[-1], it.colors[-1], dev, depth_filter=True)
        with torch.no_grad():
            inside = g <= torch.minimum(10 * g.median(), 1.2 * torch.max(g))
        keep = torch.nonzero(inside, as_tuple=True)[0]
        o, d, g, col = o[keep], d[keep], g[keep], col[keep]
        rq = None
        depth, var, color, valid = it.rend.render_batch_ray(it.npc, it.model, d, o, dev, 'color', gt_depth=g,
                                                            npc_geo_feats=it.npc_geo, npc_col_feats=it.npc_col,
                                                            is_tracker=True, cloud_pos=it.cloud, dynamic_r_query=rq)
        L.tracker_loss(depth, var, color, g, col, True, True, 0.5)[0].backward()     # src/Tracker.py:171-193
        rays[0] = o.shape[0]
        return rays[0]
    try:
        if it.dyn:
            return {'skipped': 'tracker line is measured on the fixed-radius (replica) config only'}
        ms, r = time_steps(trk_step, n, flush, dev)
        return {'rays_per_s': r / (ms * 1e-3), 'ms_per_step': ms, 'rays_per_step': r}
    finally:
        it.rend.sigmoid_coefficient = cfg['rendering']['sigmoid_coef_mapper']


# ------------------------------------------------------------------------------ CPU arms (oracle; the only place bench.py runs oracle/)
class OracleIteration:
    """The same mapping iteration restated on the CPU: oracle/sampling.py + oracle/render.py + C grid k-NN."""

    def __init__(self, dataset, n_points, stage, pixels_total=None):
        import loopy_slam_b200 as L
        from oracle import render as orc
        from oracle import sampling as osm
        from oracle.knn_c import GridKNN
        self.orc, self.osm = orc, osm
        sc = build_scene(dataset, n_points)
        self.sc = sc
        cfg = sc['cfg']
        self.cfg = cfg
        self.stage = stage
        torch.manual_seed(SEED)
        model = L.get_model(cfg)
        self.W = {k: v.detach().clone() for k, v in model.state_dict().items()}
        self.W['color_decoder.embedder._B'] = model.color_decoder.embedder._B.clone()
        self.ocfg = orc.OracleCfg.from_cfg(cfg)
        self.grid = GridKNN(sc['cloud'], float(cfg['pointcloud']['radius_query']))
        self.n_frames = len(sc['frames'])
        self.pixels_total = pixels_total or cfg['mapping']['pixels']
        self.pix_per_image = self.pixels_total // self.n_frames
        self.room = sc['room']
        self.w_color = cfg['mapping']['w_color_loss']
        if cfg['use_dynamic_radius'] or cfg['model']['encode_exposure']:
            raise NotImplementedError('the CPU arm times the replica (headline) configuration')
        torch.manual_seed(SEED)

    def step(self, outputs=None, fixed_rays=None):
        orc, osm, room, sc = self.orc, self.osm, self.room, self.sc
        t0 = time.perf_counter()
        geo = sc['geo'].clone().requires_grad_(True)                 # stands for the index_put + leaf blocks (same arithmetic)
        col = sc['col'].clone().requires_grad_(True)
        if fixed_rays is None:
            O, D, G, Cc = [], [], [], []
            for f in range(self.n_frames):
                color, depth, c2w = sc['frames'][f]
                o, d, g, c, i, j = osm.get_samples(0, room.H, 0, room.W, self.pix_per_image, room.H, room.W, room.fx, room.fy,
                                                   room.cx, room.cy, c2w, depth, color, depth_filter=True)
                O.append(o); D.append(d); G.append(g); Cc.append(c)
            o, d, g, c = torch.cat(O), torch.cat(D), torch.cat(G), torch.cat(Cc)
            inside = g <= torch.minimum(10 * g.median(), 1.2 * torch.max(g))
            o, d, g, c = o[inside], d[inside], g[inside], c[inside]
        else:
            o, d, g, c = fixed_rays
        z = orc.sample_z(g, self.ocfg)
        p = (o[:, None, :] + d[:, None, :] * z[:, :, None]).reshape(-1, 3)
        knn = self.grid.query(p, self.ocfg.radius_query)
        Wl = {k: v.clone().requires_grad_(k.startswith('color_decoder') and k != 'color_decoder.embedder._B')
              for k, v in self.W.items()}
        depth, var, rgb, valid, _ = orc.render_rays(Wl, self.ocfg, o, d, g, geo, col, sc['cloud'], self.stage, knn=knn)
        loss = mapper_loss_reference(depth, rgb, valid, g, c, self.stage, self.w_color)
        loss.backward()
        if outputs is not None:
            outputs.update(depth=depth.detach(), rgb=rgb.detach(), valid=valid.detach(), gt=g)
        return time.perf_counter() - t0, o.shape[0]


def cpu_baseline(args, it):
    """Same-run CPU baseline (rank 0, N = 1) on a bounded sample + the quality half of the metric: depth-L1 of the CUDA
    render against the oracle render of the SAME rays."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oi = OracleIteration(args.config, args.n_points, args.stage)
    # parity: render one sampled batch with both implementations
    with torch.no_grad():
        o, d, g, c, rq, _ = it.sample()
        dep, _, colr, val = it.rend.render_batch_ray(it.npc, it.model, d, o, it.dev, args.stage, gt_depth=g,
                                                     npc_geo_feats=it.npc_geo, npc_col_feats=it.npc_col, is_tracker=False,
                                                     cloud_pos=it.cloud)
    torch.cuda.synchronize(it.dev)
    ref = {}
    oi.step(outputs=ref, fixed_rays=(o.cpu(), d.cpu(), g.cpu(), c.cpu()))
    ts, rays = [], []
    t_start = time.perf_counter()
    while len(ts) < 3 or (time.perf_counter() - t_start < args.cpu_budget and len(ts) < 20):
        t, r = oi.step()
        ts.append(t); rays.append(r)
    t = sum(ts) / len(ts)
    R = sum(rays) / len(rays)
    out = {'value': R / t, 'unit': 'rays/s', 'cores': cores, 'kind': 'port',
           'sample': f'{len(ts)} full mapping iterations ({R:.0f} rays x 5 samples, N={oi.sc["cloud"].shape[0]}) on the oracle '
                     f'(oracle/sampling.py + oracle/render.py, torch CPU, + C grid k-NN), {cores} threads', 'ms_per_step': t * 1e3}
    d_cu, c_cu, v_cu = dep.cpu(), colr.cpu(), val.cpu()
    ok = ref['valid'].bool() & (ref['gt'] > 0)
    out['parity'] = {'depth_l1_vs_oracle_m': float((d_cu - ref['depth']).abs()[ok].mean()) if ok.any() else 0.0,
                     'rgb_l1_vs_oracle': float((c_cu - ref['rgb']).abs()[ok].mean()) if ok.any() else 0.0,
                     'valid_mask_identical': bool((v_cu.bool() == ref['valid'].bool()).all()), 'rays': int(ok.sum()),
                     'mean_depth_m': float(ref['depth'][ok].mean()) if ok.any() else 0.0}
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg0 = make_cfg(args.config)
    pixels_total = cfg0['mapping']['pixels']
    oi = OracleIteration(args.config, args.n_points, args.stage, pixels_total)
    for w in range(args.warmup):
        oi.step()
    ts, rays = 0.0, 0
    for k in range(args.steps):
        t, r = oi.step()
        ts += t
        rays += r
    ms = ts * 1e3 / args.steps
    val = rays / ts
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'rays/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': args.scaling,
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.config, args.stage, oi.sc['cloud'].shape[0], world, pixels_total, oi.n_frames, oi.room.H,
                                  oi.room.W, args.scaling),
        'rays_per_step': rays / args.steps,
        'cpu_baseline': {'value': val, 'unit': 'rays/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{args.steps} full mapping iterations ({rays / args.steps:.0f} rays x 5 samples each)'},
        'e2e': {'value': val, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'the reference is pure PyTorch + faiss-gpu and cannot travel to / run on this box; timed: the oracle restatement of '
                'the same iteration on the host cores (kind "port")'}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='lsr', choices=['lsr', 'reference'])
    ap.add_argument('--stage', default='color', choices=['color', 'geometry'])
    ap.add_argument('--config', default='replica', choices=['replica', 'tum', 'scannet'])
    ap.add_argument('--n-points', type=int, default=N_POINTS)
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the labelled side measurements')
    ap.add_argument('--no-sweep', action='store_true', help='skip the TUM / ScanNet / N-sweep lines of `extra`')
    ap.add_argument('--sweep-budget-s', type=float, default=150.0, help='side measurements of `extra` start only while the run is younger than this')
    ap.add_argument('--cpu-budget', type=float, default=20.0)
    args = ap.parse_args()
    if args.impl == 'reference':
        rank = int(os.environ.get('RANK', '0'))
        return run_reference(args, rank, int(os.environ.get('WORLD_SIZE', '1')))
    args.warmup = max(args.warmup, 3)
    from loopy_slam_b200 import parallel
    rank, world, local = parallel.init_from_env()
    try:
        run_lsr(args, rank, world, local)
    finally:
        if world > 1 and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
