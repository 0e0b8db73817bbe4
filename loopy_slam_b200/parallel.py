"""Multi-GPU data parallelism of the render hot path.

Rays are independent given replicated {cloud positions, feature tables, decoder weights}
(<= N*(12+256) B ~ 0.27 GB at N = 1e6), so the path shards by RAYS: one process per GPU
(torch.distributed, NCCL over NVLink/NVSwitch), every rank renders its own slice of the
iteration's rays with the fused kernels, and ONE all-reduce(sum) per optimiser step over a flat
buffer [d_geo_feats(n_sel x 32) | d_col_feats(n_sel x 32) | decoder grads | pose grad] makes the
gradients identical everywhere; every rank then applies the same Adam step, so the replicas stay
in sync without any broadcast.  The reference losses are *sums* over rays (src/Mapper.py:693-717,
src/Tracker.py:183-188), so sharded-sum + all-reduce is exact up to fp32 addition order.

Batch-global statistics that must not be sharded (the mapper's inside_mask uses the batch median,
src/Mapper.py:674-676) are computed on the full batch before slicing -- see shard_rays().
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*).
    Returns (rank, world, local_rank).  World size 1 (no env) needs no process group."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_bounds(n, rank, world):
    """Contiguous slice [lo, hi) of n rays owned by `rank` (per-frame sub-batches stay contiguous,
    needed by the per-frame exposure slices of src/Mapper.py:700-714)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(tensors, rank, world):
    """Slice every (R, ...) tensor of a ray batch to this rank's contiguous share."""
    n = tensors[0].shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return [t[lo:hi] if t is not None else None for t in tensors]


class GradAllReducer:
    """One gradient exchange per optimiser step.  Small tensors (decoder weights, pose) travel in ONE
    flat fp32 buffer; tensors above ``big_numel`` (the two trainable feature sub-blocks, ~9 MB each)
    are all-reduced in place -- staging them through the flat buffer would cost two extra full
    copies per step.  All collectives are issued back to back (async) and waited for together."""

    def __init__(self, params, group=None, big_numel=1 << 18):
        self.params = [p for p in params]
        self.group = group
        self.big_numel = big_numel
        self.small = [p for p in self.params if p.numel() < big_numel]
        self.big = [p for p in self.params if p.numel() >= big_numel]
        self.numel = sum(p.numel() for p in self.small)
        self.flat = None

    def _ensure(self, device):
        if self.flat is None or self.flat.device != device or self.flat.numel() != self.numel:
            self.flat = torch.zeros(max(self.numel, 1), dtype=torch.float32, device=device)
        return self.flat

    @staticmethod
    def _inside(t, buf):
        """True when tensor t is a view into buf's memory."""
        if t is None or buf is None or t.device != buf.device or not t.is_contiguous():
            return False
        lo = buf.data_ptr()
        return lo <= t.data_ptr() and t.data_ptr() + t.numel() * t.element_size() <= lo + buf.numel() * buf.element_size()

    def allreduce_(self, grad_buffer=None):
        """Sum .grad of all params across ranks in place (missing grads count as zero).  Returns the
        number of bytes exchanged per rank.
        grad_buffer: the flat buffer the last fused backward accumulated into (Renderer.last_grad_buffer).
        Every .grad that is a view of it (autograd keeps the views the backward returned) is covered by ONE
        all-reduce of that buffer -- no per-parameter staging copies; the others take the generic path."""
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return 0
        dev = next((p.grad.device for p in self.params if p.grad is not None), self.params[0].device)
        works, nbytes = [], 0
        covered = set()
        if grad_buffer is not None and grad_buffer.numel() > 0:
            cov = [p for p in self.params if self._inside(p.grad, grad_buffer)]
            covered = {id(p) for p in cov}
            if covered:
                # only the span of the buffer that holds .grad views travels: with the reference's index_put flow the
                # buffer also carries the two TABLE-sized feature gradients (N x 32 each), which nobody owns as .grad
                # (autograd gathers the leaf rows out of them) and which must not be exchanged
                lo = min(p.grad.data_ptr() for p in cov)
                hi = max(p.grad.data_ptr() + p.grad.numel() * 4 for p in cov)
                a = (lo - grad_buffer.data_ptr()) // 4
                b = (hi - grad_buffer.data_ptr()) // 4
                span = grad_buffer[a:b]
                works.append(dist.all_reduce(span, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
                nbytes += span.numel() * 4
        if len(covered) == len(self.params):
            for w in works:
                w.wait()
            return nbytes
        big = [p for p in self.big if id(p) not in covered]
        small = [p for p in self.small if id(p) not in covered]
        for p in big:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            works.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            if g is not p.grad:
                p.grad = g
            nbytes += g.numel() * 4
        flat = self._ensure(dev)
        off, views = 0, []
        for p in small:
            n = p.numel()
            v = flat[off:off + n]
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad.reshape(-1))
            views.append(v)
            off += n
        if small:
            works.append(dist.all_reduce(flat[:max(off, 1)], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            nbytes += off * 4
        for w in works:
            w.wait()
        for p, v in zip(small, views):
            if p.grad is None:
                p.grad = v.view(p.shape).clone()
            else:
                p.grad.copy_(v.view(p.shape))
        return nbytes


def max_over_ranks(value, device):
    """Max of a python float over ranks (device-side timing is reported as the max over ranks)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def render_img_sharded(renderer, npc, decoders, c2w, device, stage, gt_depth=None, npc_geo_feats=None, npc_col_feats=None,
                       dynamic_r_query=None, cloud_pos=None, exposure_feat=None, group=None):
    """Renderer.render_img (/root/reference/src/utils/Renderer.py:203-276: 816 k forward-only rays on Replica) with the
    ray tiles sharded over the ranks and ONE all-gather of the per-ray outputs.  Shards are whole groups of
    `ray_batch_size` rays, so the per-tile far statistic of zero-depth rays (Renderer.py:102-121) is computed on exactly
    the same rays as on one GPU: the result is bit-identical to render_img.  -> depth (H,W) f64, unc (H,W) f64, colour."""
    from .common import get_rays
    from .renderer import fused_render
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return renderer.render_img(npc, decoders, c2w, device, stage, gt_depth=gt_depth, npc_geo_feats=npc_geo_feats,
                                   npc_col_feats=npc_col_feats, dynamic_r_query=dynamic_r_query, cloud_pos=cloud_pos,
                                   exposure_feat=exposure_feat)
    with torch.no_grad():
        H, W = renderer.H, renderer.W
        n = H * W
        rbs = renderer.ray_batch_size
        groups = (n + rbs - 1) // rbs
        per = (groups + world - 1) // world                      # ray groups per rank
        lo, hi = min(rank * per * rbs, n), min((rank + 1) * per * rbs, n)
        rays_o, rays_d = get_rays(H, W, renderer.fx, renderer.fy, renderer.cx, renderer.cy, c2w, device)
        rays_o, rays_d = rays_o.reshape(-1, 3)[lo:hi], rays_d.reshape(-1, 3)[lo:hi]
        dyn = dynamic_r_query.reshape(-1)[lo:hi] if (renderer.use_dynamic_radius and dynamic_r_query is not None) else None
        gt = gt_depth.reshape(-1)[lo:hi] if gt_depth is not None else None
        out = torch.zeros(per * rbs, 5, dtype=torch.float32, device=rays_o.device)     # depth | var | rgb, padded to the shard size
        if hi > lo:
            depth, var, rgb, _ = fused_render(renderer, npc, decoders, rays_d, rays_o, stage, gt, npc_geo_feats, npc_col_feats,
                                              False, cloud_pos, dyn, exposure_feat, far_group=rbs)
            out[:hi - lo, 0], out[:hi - lo, 1], out[:hi - lo, 2:] = depth, var, rgb
        full = torch.empty(world * per * rbs, 5, dtype=torch.float32, device=out.device)
        dist.all_gather_into_tensor(full, out, group=group)
        full = full[:n]
        return full[:, 0].double().reshape(H, W), full[:, 1].double().reshape(H, W), full[:, 2:].reshape(H, W, 3).contiguous()
