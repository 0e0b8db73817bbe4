#!/usr/bin/env python
"""Headline benchmark: rendered rays/s (forward + backward) of Loopy-SLAM's per-iteration render
hot path on a synthetic Replica-room0-shaped stream (BASELINE.json metric).

One "step" = one pass of the hot path over one mapper-shaped ray batch (12 frames x 416 uniform
pixels, depth > 0  =>  ~4940 rays x 5 samples, N = 2e5 neural points, stage 'color', Replica
decoder flags):  npc_feats[indices] = leaf sub-block (src/Mapper.py:581-582)  ->
Renderer.render_batch_ray (fused sm_100a forward)  ->  mapper loss (src/Mapper.py:689-720)  ->
backward (fused sm_100a backward: feature scatter, decoder weight grads).  No optimiser step.

  value : device-timed (CUDA events), ray batch already resident in HBM.
  e2e   : the same step through the same public API, but every step's ray batch starts in PINNED
          HOST memory (H2D inside the timed region) and the loss is read back (D2H).
  --impl reference : the oracle's torch-CPU restatement of the reference path (+ exact C grid
          k-NN) on the host cores -- the reference itself is pure Python/PyTorch+FAISS-GPU and cannot
          travel to / run on this box, see DESIGN.md.
Multi-GPU (torchrun, one rank per GPU): weak scaling -- every rank renders its own full batch, then
ONE NCCL all-reduce over [d_geo_sub | d_col_sub | decoder grads] per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_PER_FRAME, N_FRAMES, N_POINTS, SEED = 416, 12, 200000, 1219
S, K, C = 5, 8, 32
BYTES_RAY_FWD = {'color': 53 + S * (K * 12 + 2 * K * C * 4), 'geometry': 53 + S * (K * 12 + 1 * K * C * 4)}
BYTES_RAY_ALL = {'color': 53 + S * (K * 12 + 2 * K * C * 4 * 3), 'geometry': 53 + S * (K * 12 + 1 * K * C * 4 * 3)}
MFLOP_RAY_FWD = {'color': 1.99, 'geometry': 0.157}          # SURVEY.md 8d (Replica flags), 2*MAC


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)', d
    return 6650.0, 'fallback (B200_PROFILING.md)', {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region: one streaming `nvidia-smi -lms`
    subprocess (started before warm-up so it is already sampling), rows kept between start()/stop()."""

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


def build_scene(n_points, cache=True):
    """Synthetic stream + cloud on the host (cached in /tmp so the reference arm and the scaling
    runs on one box do not rebuild it)."""
    from loopy_slam_b200.stream import SyntheticRoom, build_point_cloud, sample_batch
    room = SyntheticRoom(seed=SEED)
    path = f'/tmp/lsr_bench_scene_{n_points}.pt'
    if cache and os.path.exists(path):
        return room, torch.load(path)
    cloud, geo, col = build_point_cloud(room, n_points, frame_stride=40, pixels_per_frame=20000, seed=SEED)
    frames = list(range(0, 40 * N_FRAMES, 40))
    batches = [sample_batch(room, frames, R_PER_FRAME, seed=SEED + 17 * b) for b in range(8)]
    cur_color, cur_depth, cur_c2w = room.frame(frames[-1])
    sc = dict(cloud=cloud, geo=geo, col=col, batches=batches, cur_depth=cur_depth, cur_c2w=cur_c2w)
    if cache:
        try:
            torch.save(sc, path + f'.{os.getpid()}')
            os.replace(path + f'.{os.getpid()}', path)
        except Exception:
            pass
    return room, sc


def mapper_loss_eager(depth, color, valid, gt_depth, gt_color, stage, w_color=0.1):
    """src/Mapper.py:689-720 as plain torch ops (kept for A/B: --eager-loss); the default step uses the fused
    lsr_mapper_loss kernel through loopy_slam_b200.mapper_loss."""
    m = ((gt_depth > 0) & valid & (~torch.isnan(depth))).to(depth.dtype)
    loss = (torch.abs(gt_depth - depth) * m).sum()
    if stage == 'color':
        loss = loss + w_color * (torch.abs(gt_color - color) * m[:, None]).sum()
    return loss


def run_lsr(args, rank, world, local):
    import loopy_slam_b200 as L
    from loopy_slam_b200 import parallel, _lib
    from loopy_slam_b200.frustum import get_mask_from_c2w
    import __graft_entry__ as entry
    if not os.path.exists(_lib.LIB_PATH):
        entry.build()
    dev = torch.device(f'cuda:{local}')
    torch.cuda.set_device(dev)
    stage = args.stage
    room, sc = build_scene(args.n_points)
    cfg = L.default_cfg('replica')
    torch.manual_seed(SEED)
    model = L.get_model(cfg).to(dev)

    class Slam:
        H, W, fx, fy, cx, cy = room.H, room.W, room.fx, room.fy, room.cx, room.cy
    rend = L.Renderer(cfg, None, Slam)
    rend.sigmoid_coefficient = cfg['rendering']['sigmoid_coef_mapper']

    class NPC:
        def get_radius_query(self):
            return cfg['pointcloud']['radius_query']
    npc = NPC()
    cloud = sc['cloud'].to(dev)
    npc_geo, npc_col = sc['geo'].to(dev), sc['col'].to(dev)
    indices = get_mask_from_c2w(cloud, sc['cur_c2w'], sc['cur_depth'].to(dev), room.H, room.W, room.fx, room.fy,
                                room.cx, room.cy, edge=-4)
    geo_leaf = npc_geo[indices].clone().requires_grad_(True)      # src/Mapper.py:502-505
    col_leaf = npc_col[indices].clone().requires_grad_(True)
    # fix_geo_decoder: True  =>  colour decoder + the two geo Fourier matrices train (src/Mapper.py:524-541)
    for p in model.geo_decoder.parameters():
        p.requires_grad_(False)
    model.geo_decoder.embedder._B.requires_grad_(True)
    train_params = [p for p in model.parameters() if p.requires_grad] + [geo_leaf, col_leaf]
    reducer = parallel.GradAllReducer(train_params)
    subset = L.FeatureSubset(indices, npc_geo.shape[0])      # built once per mapped frame, like `indices` itself
    dev_batches = [[t.to(dev) for t in b] for b in sc['batches']]
    # e2e inputs: one pinned staging buffer per batch [rays_o | rays_d | depth | colour] -> ONE H2D copy per step
    def pack_host(b):
        flat = torch.cat([t.reshape(-1).to(torch.float32) for t in b]).pin_memory()
        return flat, [tuple(t.shape) for t in b]
    host_batches = [pack_host(b) for b in sc['batches']]

    def h2d(hb):
        flat, shapes = hb
        d = flat.to(dev, non_blocking=True)
        out, off = [], 0
        for shp in shapes:
            n = 1
            for v in shp:
                n *= v
            out.append(d[off:off + n].view(shp))
            off += n
        return out
    R = dev_batches[0][0].shape[0]
    timing = {'fwd': [], 'bwd': []}
    launches = [0]

    def step(batch, timed=False):
        o, d, g, c = batch
        for p in train_params:
            p.grad = None
        rend._timing = timing if timed else None
        if args.index_put:      # the reference's literal flow: table[indices] = leaf every iteration (src/Mapper.py:581-582)
            gtab = npc_geo.index_put((indices,), geo_leaf)
            ctab = npc_col.index_put((indices,), col_leaf)
            depth, var, color, valid = rend.render_batch_ray(npc, model, d, o, dev, stage, gt_depth=g,
                                                             npc_geo_feats=gtab, npc_col_feats=ctab, is_tracker=False,
                                                             cloud_pos=cloud)
        else:                   # same sub-block, read through lsr's row_remap (no table rewrite, leaf-sized gradients)
            depth, var, color, valid = rend.render_batch_ray(npc, model, d, o, dev, stage, gt_depth=g,
                                                             npc_geo_feats=npc_geo, npc_col_feats=npc_col,
                                                             is_tracker=False, cloud_pos=cloud,
                                                             feat_subset=(subset, geo_leaf, col_leaf))
        if args.eager_loss:
            loss = mapper_loss_eager(depth, color, valid, g, c, stage)
        else:
            loss = L.mapper_loss(depth, color, valid, g, c, stage, 0.1)[0]     # fused lsr_mapper_loss (src/Mapper.py:689-720)
        loss.backward()
        rend._timing = None
        if world > 1:
            reducer.allreduce_(getattr(rend, 'last_grad_buffer', None))
        launches[0] += 6            # far_bound + weight re-layout + sample_knn + render_fwd + mapper_loss + render_bwd (ours); torch ops not counted
        return loss

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local) if rank == 0 else None
    for w in range(args.warmup):
        step(dev_batches[w % len(dev_batches)])
    barrier()
    if sampler:
        sampler.start()
    launches[0] = 0
    evs = []
    for k in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(dev_batches[k % len(dev_batches)], timed=True)
        e1.record()
        evs.append((e0, e1))
    barrier()
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    ms_step = parallel.max_over_ranks(ms_total / args.steps, dev)
    n_launch = launches[0]
    kt = {k: [a.elapsed_time(b) for a, b in v] for k, v in timing.items()}
    # ---- e2e: host buffers in, loss out, wall-clock bracketed by syncs
    for w in range(2):
        step(h2d(host_batches[w])).item()
    barrier()
    prof = None
    if args.host_profile:      # where the host time of the public call goes (stderr; not part of the JSON line)
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    t0 = time.perf_counter()
    for k in range(args.steps):
        hb = host_batches[k % len(host_batches)]
        loss = step(h2d(hb))
        loss_host = loss.item()                                        # D2H of the step's result
    barrier()
    if prof is not None:
        import pstats
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats('tottime').print_stats(28)
    e2e_ms = parallel.max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps, dev)
    clocks = sampler.stop() if sampler else None
    h2d_bytes = host_batches[0][0].numel() * host_batches[0][0].element_size()

    extra = {}
    if world == 1 and not args.no_extra:
        def timed(fn, n):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(dev)
            evs2 = []
            for k in range(n):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(k); e1.record()
                evs2.append((e0, e1))
            torch.cuda.synchronize(dev)
            return sum(a.elapsed_time(b) for a, b in evs2) / n
        n_extra = max(10, args.steps // 4)
        # (1) mapper iteration, stage 'geometry' (40 % of the mapping iterations, src/Mapper.py:588-591)
        stage_saved = stage

        def geo_step(k=0):
            o, d, g, c = dev_batches[k % len(dev_batches)]
            for p in train_params:
                p.grad = None
            depth, var, color, valid = rend.render_batch_ray(npc, model, d, o, dev, 'geometry', gt_depth=g,
                                                             npc_geo_feats=npc_geo, npc_col_feats=npc_col,
                                                             is_tracker=False, cloud_pos=cloud,
                                                             feat_subset=(subset, geo_leaf, col_leaf))
            L.mapper_loss(depth, color, valid, g, c, 'geometry', 0.1)[0].backward()
        ms = timed(geo_step, n_extra)
        extra['mapper_geometry_stage'] = {'rays_per_s': R / (ms * 1e-3), 'ms_per_step': ms, 'rays': R}
        # (2) tracking iteration (src/Tracker.py:102-197): 1500 rays, stage 'color', is_tracker, frozen decoders,
        #     gradient to the ray origins/directions (-> pose)
        for p in model.parameters():
            p.requires_grad_(False)
        tr_batches = [[t[:1500].clone() for t in b] for b in dev_batches]

        def trk_step(k=0):
            o, d, g, c = tr_batches[k % len(tr_batches)]
            o = o.detach().requires_grad_(True)
            d = d.detach().requires_grad_(True)
            depth, var, color, valid = rend.render_batch_ray(npc, model, d, o, dev, 'color', gt_depth=g,
                                                             npc_geo_feats=npc_geo, npc_col_feats=npc_col,
                                                             is_tracker=True, cloud_pos=cloud)
            L.tracker_loss(depth, var, color, g, c, True, True, 0.5)[0].backward()     # src/Tracker.py:171-193
        ms = timed(trk_step, n_extra)
        extra['tracker_iteration'] = {'rays_per_s': 1500 / (ms * 1e-3), 'ms_per_step': ms, 'rays': 1500}

    if rank != 0:
        return
    peak, peak_src, peaks = load_peaks()
    mean = lambda xs: sum(xs) / max(len(xs), 1)
    t_f, t_b = mean(kt['fwd']), mean(kt['bwd'])
    dom = 'render_bwd_kernel' if t_b >= t_f else 'render_fwd_kernel'
    t_dom = max(t_b, t_f)
    bytes_dom = R * ((BYTES_RAY_ALL[stage] - BYTES_RAY_FWD[stage]) if t_b >= t_f else BYTES_RAY_FWD[stage])
    ach = bytes_dom / (t_dom * 1e-3) / 1e9 if t_dom > 0 else 0.0
    sm_mhz = (clocks or {}).get('sm_mhz') or 1965
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    flops_step = R * MFLOP_RAY_FWD[stage] * 1e6 * 3
    out = {
        'metric': 'rendered rays/sec (fwd+bwd)', 'value': world * R / (ms_step * 1e-3), 'unit': 'rays/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f"Replica room0 default config, mapping iteration hot path, stage '{stage}', "
                               f"Replica decoder flags (rel-pos neighbour MLP on)",
                   'rays_per_step_per_gpu': R, 'n_surface': S, 'n_points': int(cloud.shape[0]),
                   'n_trainable_rows': int(indices.shape[0]), 'image': f'{room.H}x{room.W}', 'frames_per_batch': N_FRAMES,
                   'parallelism': f'ray-shard dp{world} (replicated cloud+weights, 1 NCCL all-reduce/step)',
                   'feature_subblock': 'index_put into the tables every step (src/Mapper.py:581-582)' if args.index_put
                                       else 'lsr row_remap + leaf blocks (same rows, no table rewrite)',
                   'loss': 'torch ops' if args.eager_loss else 'lsr_mapper_loss',
                   'l2': 'flushed between timed steps (256 MiB memset outside the event pairs)'},
        'e2e': {'value': world * R / (e2e_ms * 1e-3), 'unit': 'rays/s', 'ms_per_step': e2e_ms,
                'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4},
        'gpu_launches': n_launch,
        'roofline': {'bound': 'hbm', 'kernel': dom, 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
                     'frac': ach / peak,
                     'traffic': (304.7e6 if t_b >= t_f else 254.5e6) if stage == 'color' and R == 4936 else None, 'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': bytes_dom,
                     'traffic_source': 'dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/r01_ncu_v10_summary.md '
                                       '(same command; the excess over the algorithmic bytes is the saved-activation round trip)',
                     'note': 'HBM fraction as defined in BASELINE.md section 3; arithmetic intensity ~190 FLOP/B puts both '
                             'fused kernels on the tensor/issue side of the roofline, see "tensor"'},
        'tensor': {'algorithmic_tflops': flops_step / ((t_f + t_b) * 1e-3) / 1e12 if t_f + t_b > 0 else 0,
                   'executed_tflops_3xtf32': 3 * flops_step / ((t_f + t_b) * 1e-3) / 1e12 if t_f + t_b > 0 else 0,
                   'peak_tf32_mma_sync_tflops_measured': 278.0, 'peak_bf16_tflops_measured': peaks.get('bf16_tflops'),
                   'forward': 'tcgen05.mma kind::tf32 (M=128, accumulators + hidden activations in TMEM, weights by '
                              'cp.async.bulk), error-compensated 3xTF32',
                   'backward': 'mma.sync m16n8k8 tf32 (HMMA.1688.F32.TF32), error-compensated 3xTF32',
                   'note': '3 tensor-core passes per fp32 product (hi*hi + lo*hi + hi*lo) keep the 1e-4 parity contract; 278 '
                           'TFLOP/s is the measured mma.sync TF32 issue peak on this B200 (tools/mma_rate.cu), the tcgen05 '
                           'tf32 rate measured in tools/umma_probe.cu is 92 cycles per 128x128x8 MMA per SM'},
        'kernels': {'render_fwd_ms': t_f, 'render_bwd_ms': t_b,
                    'fwd_bwd_GBps': R * BYTES_RAY_ALL[stage] / ((t_f + t_b) * 1e-3) / 1e9 if t_f + t_b > 0 else 0},
        'fp32_fma_peak_tflops_at_clock': fp32_peak,
        'extra': extra,
        'clocks': clocks, 'loss': loss_host,
    }
    if world == 1 and not args.no_cpu_baseline:
        for p in model.parameters():
            p.requires_grad_(False)
        o0, d0, g0, _ = dev_batches[0]
        with torch.no_grad():
            dep, _, colr, val = rend.render_batch_ray(npc, model, d0, o0, dev, stage, gt_depth=g0, npc_geo_feats=npc_geo,
                                                      npc_col_feats=npc_col, is_tracker=False, cloud_pos=cloud)
        torch.cuda.synchronize(dev)
        out['cpu_baseline'] = cpu_baseline(sc, stage, budget_s=args.cpu_budget, cuda_render=(dep, colr, val))
        out['depth_l1_vs_reference_m'] = out['cpu_baseline']['parity']['depth_l1_vs_oracle_m']
    print(json.dumps(out))


# ------------------------------------------------------------------------------ CPU arms
def _oracle_setup(sc, stage):
    import loopy_slam_b200 as L
    from oracle import render as orc
    from oracle.knn_c import GridKNN
    cfg = L.default_cfg('replica')
    torch.manual_seed(SEED)
    model = L.get_model(cfg)
    W = {k: v.detach().clone() for k, v in model.state_dict().items()}
    W['color_decoder.embedder._B'] = model.color_decoder.embedder._B.clone()
    ocfg = orc.OracleCfg.from_cfg(cfg)
    grid = GridKNN(sc['cloud'], 0.08)
    return orc, ocfg, W, grid


def _oracle_step(orc, ocfg, W, grid, sc, batch, stage, outputs=None):
    """The reference's step restated on the CPU: exact 8-NN + decoders + compositing + loss + backward."""
    o, d, g, c = batch
    t0 = time.perf_counter()
    z = orc.sample_z(g, ocfg)
    p = (o[:, None, :] + d[:, None, :] * z[:, :, None]).reshape(-1, 3)
    knn = grid.query(p, ocfg.radius_query)
    Wl = {k: v.clone().requires_grad_(k.startswith('color_decoder') and k != 'color_decoder.embedder._B')
          for k, v in W.items()}
    geo = sc['geo'].clone().requires_grad_(True)
    col = sc['col'].clone().requires_grad_(True)
    depth, var, rgb, valid, _ = orc.render_rays(Wl, ocfg, o, d, g, geo, col, sc['cloud'], stage, knn=knn)
    loss = mapper_loss_eager(depth, rgb, valid, g, c, stage)
    loss.backward()
    if outputs is not None:
        outputs.update(depth=depth.detach(), rgb=rgb.detach(), valid=valid.detach())
    return time.perf_counter() - t0


def cpu_baseline(sc, stage, budget_s=20.0, rays=None, cuda_render=None):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc, ocfg, W, grid = _oracle_setup(sc, stage)
    batch = sc['batches'][0]
    if rays:
        batch = [t[:rays] for t in batch]
    R = batch[0].shape[0]
    ref = {}
    _oracle_step(orc, ocfg, W, grid, sc, batch, stage, outputs=ref)     # warm-up; its render is the parity reference
    ts = []
    t_start = time.perf_counter()
    while len(ts) < 3 or (time.perf_counter() - t_start < budget_s and len(ts) < 20):
        ts.append(_oracle_step(orc, ocfg, W, grid, sc, batch, stage))
    t = sum(ts) / len(ts)
    out = {'value': R / t, 'unit': 'rays/s', 'cores': cores, 'kind': 'port',
           'sample': f'{len(ts)} iterations of the same step ({R} rays x 5 samples, N={sc["cloud"].shape[0]}) on the '
                     f'oracle torch-CPU restatement + C grid k-NN, {cores} threads', 'ms_per_step': t * 1e3}
    if cuda_render is not None:
        # second half of BASELINE's metric: depth-L1 of the CUDA render against the reference (oracle) render of the
        # same rays (the reference's depth_l1_render, src/Mapper.py:1146-1147, restricted to this batch)
        d_cu, c_cu, v_cu = [t.detach().cpu() for t in cuda_render]
        ok = ref['valid'].bool() & (batch[2] > 0)
        same_mask = bool((v_cu.bool() == ref['valid'].bool()).all())
        dl1 = float((d_cu - ref['depth']).abs()[ok].mean()) if ok.any() else 0.0
        cl1 = float((c_cu - ref['rgb']).abs()[ok].mean()) if ok.any() else 0.0
        out['parity'] = {'depth_l1_vs_oracle_m': dl1, 'rgb_l1_vs_oracle': cl1, 'valid_mask_identical': same_mask,
                         'rays': int(ok.sum()), 'mean_depth_m': float(ref['depth'][ok].mean()) if ok.any() else 0.0,
                         'depth_l1_vs_sensor_m_cuda': float((d_cu - batch[2]).abs()[ok].mean()) if ok.any() else 0.0,
                         'depth_l1_vs_sensor_m_oracle': float((ref['depth'] - batch[2]).abs()[ok].mean()) if ok.any() else 0.0}
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    room, sc = build_scene(args.n_points)
    stage = args.stage
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc, ocfg, W, grid = _oracle_setup(sc, stage)
    batches = sc['batches']                                             # one full mapper batch per step (~3 s of CPU work)
    R = batches[0][0].shape[0]
    for w in range(args.warmup):
        _oracle_step(orc, ocfg, W, grid, sc, batches[w % 8], stage)
    t0 = time.perf_counter()
    for k in range(args.steps):
        _oracle_step(orc, ocfg, W, grid, sc, batches[k % 8], stage)
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    val = R / (ms * 1e-3)
    print(json.dumps({
        'impl': 'reference', 'metric': 'rendered rays/sec (fwd+bwd)', 'value': val, 'unit': 'rays/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f"Replica room0 default config, mapping iteration hot path, stage '{stage}', Replica "
                               f"decoder flags (rel-pos neighbour MLP on)", 'rays_per_step': R, 'n_surface': S,
                   'n_points': int(sc['cloud'].shape[0]),
                   'note': 'reference = pure PyTorch + faiss-gpu, cannot travel; timed: oracle torch-CPU restatement of '
                           'its math + exact C grid k-NN on the host cores'},
        'cpu_baseline': {'value': val, 'unit': 'rays/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{R} rays x 5 samples per step, N={sc["cloud"].shape[0]}'},
        'e2e': {'value': val, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='lsr', choices=['lsr', 'reference'])
    ap.add_argument('--stage', default='color', choices=['color', 'geometry'])
    ap.add_argument('--n-points', type=int, default=N_POINTS)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--index-put', action='store_true',
                    help='A/B: rewrite the feature tables every step (src/Mapper.py:581-582) instead of lsr row_remap')
    ap.add_argument('--eager-loss', action='store_true', help='A/B: the mapper loss as ~30 torch ops instead of lsr_mapper_loss')
    ap.add_argument('--no-extra', action='store_true', help='skip the geometry-stage / tracker side measurements')
    ap.add_argument('--cpu-budget', type=float, default=20.0)
    ap.add_argument('--host-profile', action='store_true', help='cProfile the e2e loop (host-side overhead of the public API)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        args.steps, args.warmup = min(args.steps, 20), min(args.warmup, 3)    # each step = seconds of CPU work
        rank = int(os.environ.get('RANK', '0'))
        return run_reference(args, rank, int(os.environ.get('WORLD_SIZE', '1')))
    from loopy_slam_b200 import parallel
    rank, world, local = parallel.init_from_env()
    try:
        run_lsr(args, rank, world, local)
    finally:
        if world > 1 and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
