"""Host-side mirror of the reference decoder surface (``NICER`` and its two MLPs).

Same constructor signature, sub-module / parameter names and shapes as
/root/reference/src/conv_onet/models/decoder.py:549-571 (NICER), :106-178 (MLP_geometry),
:345-429 (MLP_color), :307-342 (MLP_col_neighbor, MLP_exposure), :12-32 (Fourier transform), so

* ``state_dict()`` / ``load_state_dict()`` interoperate with the reference's checkpoints and the
  pretrained ``middle_fine.pt`` (src/Point_SLAM.py:185-198, src/utils/Logger.py:33),
* the optimiser groups the mapper builds from ``decoders.*.parameters()`` (src/Mapper.py:524-541)
  see the same Parameters.

The arithmetic does NOT live here: the modules are parameter containers.  All parameters are views
into one flat, 16-byte-aligned fp32 device blob (``WeightBlob``) that the fused CUDA kernels read
directly through the ``LsrWeights`` offset table (include/lsr.h).
"""
import math

import torch
import torch.nn as nn

from . import _lib

C_DIM, H_GEO, H_COL, E_GEO, E_COL, E_REL = 32, 32, 128, 93, 20, 10


class GaussianFourierFeatureTransform(nn.Module):
    """Parameter holder for the random Fourier matrix ``_B`` (decoder.py:22-32).  Non-learnable
    transforms keep ``_B`` as a plain tensor attribute exactly like the reference (so it is neither
    in ``parameters()`` nor in ``state_dict()``)."""

    def __init__(self, num_input_channels, mapping_size=93, scale=25, learnable=False, concat=True):
        super().__init__()
        self.concat, self.mapping_size, self.scale, self.learnable = concat, mapping_size, scale, learnable
        B = torch.randn((num_input_channels, mapping_size)) * scale
        if learnable:
            self._B = nn.Parameter(B)
        else:
            self._B = B

    def forward(self, x):   # kept for API completeness (tools / debugging); not on the fused path
        x = x.squeeze(0)
        y = (2 * math.pi * x) @ self._B.to(x.device)
        return torch.cat((torch.sin(y), torch.cos(y)), -1) if self.concat else torch.sin(y)


class DenseLayer(nn.Linear):
    """nn.Linear with xavier-uniform(gain(activation)) weight and zero bias (decoder.py:84-93)."""

    def __init__(self, in_dim, out_dim, activation='relu', *args, **kwargs):
        self.activation = activation
        super().__init__(in_dim, out_dim, *args, **kwargs)

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.weight, gain=nn.init.calculate_gain(self.activation))
        if self.bias is not None:
            nn.init.zeros_(self.bias)


class MLP_col_neighbor(nn.Module):
    def __init__(self, c_dim, embedding_size_rel, hidden_size):
        super().__init__()
        self.linear1 = nn.Linear(c_dim + embedding_size_rel, hidden_size)
        self.linear2 = nn.Linear(hidden_size, c_dim)
        nn.init.xavier_uniform_(self.linear1.weight)
        nn.init.xavier_uniform_(self.linear2.weight)


class MLP_exposure(nn.Module):
    """8 -> 128 -> 12 affine-colour MLP (decoder.py:326-342).  Tiny and evaluated once per call
    (not per sample), so it stays in PyTorch; its 12 outputs feed the fused kernel."""

    def __init__(self, latent_dim, hidden_size):
        super().__init__()
        self.linear1 = nn.Linear(latent_dim, hidden_size)
        self.linear2 = nn.Linear(hidden_size, 12)
        self.act_fn = nn.Softplus(beta=100)
        nn.init.normal_(self.linear1.weight, mean=0, std=0.01)
        nn.init.normal_(self.linear2.weight, mean=0, std=0.01)

    def forward(self, x):
        return self.linear2(self.act_fn(self.linear1(x)))


class MLP_geometry(nn.Module):
    """Parameter container of the occupancy decoder (decoder.py:125-178): hidden 32, Fourier 93."""

    def __init__(self, cfg, name='', dim=3, c_dim=32, hidden_size=128, n_blocks=5, leaky=False,
                 sample_mode='bilinear', color=False, skips=[2], pos_embedding_method='fourier',
                 concat_feature=False, use_view_direction=False):
        super().__init__()
        if use_view_direction:
            raise NotImplementedError('use_view_direction is off in every shipped config')
        self.name, self.color, self.c_dim, self.n_blocks, self.skips = name, color, c_dim, n_blocks, skips
        self.hidden_size = hidden_size
        self.fc_c = nn.ModuleList([nn.Linear(c_dim, hidden_size) for _ in range(n_blocks)])
        self.embedder = GaussianFourierFeatureTransform(dim, mapping_size=E_GEO, scale=25, concat=False, learnable=True)
        self.embedder_rel_pos = GaussianFourierFeatureTransform(3, mapping_size=E_REL, scale=32, learnable=True)
        self.mlp_col_neighbor = MLP_col_neighbor(c_dim, 2 * E_REL, hidden_size)   # unused by the reference too
        emb = E_GEO
        self.pts_linears = nn.ModuleList(
            [DenseLayer(emb, hidden_size, activation='relu')] +
            [DenseLayer(hidden_size, hidden_size, activation='relu') if i not in skips
             else DenseLayer(hidden_size + emb, hidden_size, activation='relu') for i in range(n_blocks - 1)])
        self.output_linear = DenseLayer(hidden_size, 1, activation='relu')


class MLP_color(nn.Module):
    """Parameter container of the colour decoder (decoder.py:364-429): hidden 128, Fourier 2x20."""

    def __init__(self, cfg, name='', dim=3, c_dim=32, hidden_size=128, n_blocks=5, leaky=False,
                 sample_mode='bilinear', color=True, skips=[2], pos_embedding_method='fourier',
                 concat_feature=False, use_view_direction=False):
        super().__init__()
        if use_view_direction:
            raise NotImplementedError('use_view_direction is off in every shipped config')
        self.name, self.color, self.c_dim, self.n_blocks, self.skips = name, color, c_dim, n_blocks, skips
        self.hidden_size = hidden_size
        self.encode_rel_pos_in_col = cfg['model']['encode_rel_pos_in_col']
        self.encode_exposure = cfg['model']['encode_exposure']
        self.fc_c = nn.ModuleList([nn.Linear(c_dim, hidden_size) for _ in range(n_blocks)])
        self.embedder = GaussianFourierFeatureTransform(dim, mapping_size=E_COL, scale=32)
        self.embedder_rel_pos = GaussianFourierFeatureTransform(3, mapping_size=E_REL, scale=32, learnable=True)
        self.mlp_col_neighbor = MLP_col_neighbor(c_dim, 2 * E_REL, hidden_size)
        if self.encode_exposure:
            self.mlp_exposure = MLP_exposure(cfg['model']['exposure_dim'], hidden_size)
        emb = 2 * E_COL
        self.pts_linears = nn.ModuleList(
            [DenseLayer(emb, hidden_size, activation='relu')] +
            [DenseLayer(hidden_size, hidden_size, activation='relu') if i not in skips
             else DenseLayer(hidden_size + emb, hidden_size, activation='relu') for i in range(n_blocks - 1)])
        self.output_linear = DenseLayer(hidden_size, 3, activation='linear')


class WeightBlob:
    """One flat fp32 device buffer holding every decoder tensor the kernels read, each at a
    4-element-aligned offset; the nn.Parameters are re-pointed to views of it (the trick DDP / apex
    use for flat buffers), so optimiser steps update the blob in place and no per-call packing is
    needed.  ``ensure()`` re-validates the aliasing cheaply on every render call and rebuilds after
    ``.to(device)`` / ``load_state_dict`` replaced storages."""

    FIELDS = None   # filled below

    def __init__(self, nicer):
        self.nicer = nicer
        g, c = nicer.geo_decoder, nicer.color_decoder
        ent = []   # (struct field, index or None, owner getter)
        for i in range(5):
            ent.append(('g_fc_w', i, g.fc_c[i], 'weight')); ent.append(('g_fc_b', i, g.fc_c[i], 'bias'))
        ent.append(('g_B', None, g.embedder, '_B'))
        for i in range(5):
            ent.append(('g_lin_w', i, g.pts_linears[i], 'weight')); ent.append(('g_lin_b', i, g.pts_linears[i], 'bias'))
        ent.append(('g_out_w', None, g.output_linear, 'weight')); ent.append(('g_out_b', None, g.output_linear, 'bias'))
        for i in range(5):
            ent.append(('c_fc_w', i, c.fc_c[i], 'weight')); ent.append(('c_fc_b', i, c.fc_c[i], 'bias'))
        ent.append(('c_B', None, c.embedder, '_B'))
        ent.append(('c_Brel', None, c.embedder_rel_pos, '_B'))
        ent.append(('c_nb1_w', None, c.mlp_col_neighbor.linear1, 'weight')); ent.append(('c_nb1_b', None, c.mlp_col_neighbor.linear1, 'bias'))
        ent.append(('c_nb2_w', None, c.mlp_col_neighbor.linear2, 'weight')); ent.append(('c_nb2_b', None, c.mlp_col_neighbor.linear2, 'bias'))
        for i in range(5):
            ent.append(('c_lin_w', i, c.pts_linears[i], 'weight')); ent.append(('c_lin_b', i, c.pts_linears[i], 'bias'))
        ent.append(('c_out_w', None, c.output_linear, 'weight')); ent.append(('c_out_b', None, c.output_linear, 'bias'))
        self.entries = ent
        off = 0
        self.offsets, self.numels = [], []
        for (_, _, owner, attr) in ent:
            n = getattr(owner, attr).numel()
            self.offsets.append(off)
            self.numels.append(n)
            off += (n + 3) // 4 * 4
        self.n_elems = off
        self.shapes = None      # filled by the renderer on first use (parameter shapes / strides, for the gradient views)
        self.strides = None
        self.flat = None
        self.struct = None
        self.geo_w_idx = [k for k, e in enumerate(ent) if e[0].startswith('g_') and e[0] != 'g_B']
        self.geo_b_idx = [k for k, e in enumerate(ent) if e[0] == 'g_B']
        self.col_w_idx = [k for k, e in enumerate(ent) if e[0].startswith('c_') and e[0] != 'c_B']

    def tensors(self):
        # nn.Module.__getattr__ is a slow path (called per step for ~45 tensors): look parameters up directly
        out = []
        for (_, _, owner, attr) in self.entries:
            t = owner._parameters.get(attr)
            out.append(t if t is not None else getattr(owner, attr))
        return out

    def _aliased(self, device):
        if self.flat is None or self.flat.device != device:
            return False
        base = self.flat.data_ptr()
        for t, off in zip(self.tensors(), self.offsets):
            if t.data_ptr() != base + 4 * off:
                return False
        return True

    def _adopt(self, device):
        """The tensors may ALREADY be views of one flat buffer that this object does not know about: the module was
        pickled to another process (NICER.__getstate__ drops the blob; torch re-creates the parameters as views of the
        one shared / CUDA-IPC storage: src/Point_SLAM.py shares `shared_decoders` between the tracker and mapper
        processes) or deep-copied.  Re-allocating would silently detach this process from the shared weights, so the
        existing storage is adopted when every tensor sits at base + 4 * offset inside it."""
        ts = self.tensors()
        t0 = ts[0]
        if t0.device != device or t0.dtype != torch.float32:
            return None
        st = t0.untyped_storage()
        base = t0.data_ptr() - 4 * self.offsets[0]
        if base < st.data_ptr() or base + 4 * self.n_elems > st.data_ptr() + st.nbytes():
            return None
        for t, off in zip(ts, self.offsets):
            if (t.device != device or t.dtype != torch.float32 or not t.is_contiguous()
                    or t.untyped_storage().data_ptr() != st.data_ptr() or t.data_ptr() != base + 4 * off):
                return None
        flat = torch.empty(0, dtype=torch.float32, device=device)
        flat.set_(st, (base - st.data_ptr()) // 4, (self.n_elems,), (1,))
        return flat

    def ensure(self, device):
        """Flat blob on ``device`` with all tensors aliased into it -> (flat, LsrWeights)."""
        device = torch.device(device)
        if device.type == 'cuda' and device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        if not self._aliased(device):
            flat = self._adopt(device) if self.flat is None else None
            if flat is None:
                flat = torch.zeros(self.n_elems, dtype=torch.float32, device=device)
                for (_, _, owner, attr), off, n in zip(self.entries, self.offsets, self.numels):
                    t = getattr(owner, attr)
                    view = flat[off:off + n].view(t.shape)
                    with torch.no_grad():
                        view.copy_(t.detach().to(device=device, dtype=torch.float32))
                    if isinstance(t, nn.Parameter):
                        t.data = view
                    else:
                        setattr(owner, attr, view)
            self.flat = flat
            W = _lib.LsrWeights()
            W.blob = flat.data_ptr()
            W.n_elems = self.n_elems
            for (field, idx, _, _), off in zip(self.entries, self.offsets):
                if idx is None:
                    setattr(W, field, off)
                else:
                    getattr(W, field)[idx] = off
            self.struct = W
        return self.flat, self.struct


class NICER(nn.Module):
    """Drop-in for the reference ``NICER`` (decoder.py:549-626): identical constructor, attribute
    names and parameter shapes.  ``forward`` decodes raw (occupancy, colour) per point like the
    reference's (used by ``Renderer.eval_points``); training goes through the fused
    ``Renderer.render_batch_ray``."""

    def __init__(self, cfg, dim=3, c_dim=32, hidden_size=128, pos_embedding_method='fourier',
                 use_view_direction=False):
        super().__init__()
        if c_dim != C_DIM or hidden_size != H_COL or dim != 3:
            raise NotImplementedError('lsr kernels are specialised for c_dim=32, hidden_size=128, dim=3')
        self.cfg_flags = dict(
            encode_rel_pos_in_col=bool(cfg['model']['encode_rel_pos_in_col']),
            encode_exposure=bool(cfg['model']['encode_exposure']),
            min_nn_num=int(cfg['pointcloud']['min_nn_num']),
            nn_num=int(cfg['pointcloud']['nn_num']),
            N_surface=int(cfg['rendering']['N_surface']),
            use_dynamic_radius=bool(cfg['use_dynamic_radius']),
            nn_weighting=cfg['pointcloud']['nn_weighting'])
        if self.cfg_flags['nn_weighting'] != 'distance':
            raise NotImplementedError("only pointcloud.nn_weighting == 'distance' (every shipped config)")
        self.geo_decoder = MLP_geometry(name='geometry', cfg=cfg, dim=dim, c_dim=c_dim, color=False, skips=[2],
                                        n_blocks=5, hidden_size=H_GEO, pos_embedding_method=pos_embedding_method)
        self.color_decoder = MLP_color(name='color', cfg=cfg, dim=dim, c_dim=c_dim, color=True, skips=[2],
                                       n_blocks=5, hidden_size=hidden_size,
                                       pos_embedding_method=pos_embedding_method,
                                       use_view_direction=cfg['use_view_direction'])
        self._blob = None

    @property
    def blob(self):
        if self._blob is None:
            self._blob = WeightBlob(self)
        return self._blob

    def share_memory(self):
        """src/Point_SLAM.py:92: `self.shared_decoders.share_memory()` before the tracker / mapper processes are
        spawned.  The flat blob is built FIRST so that every parameter is a view of one storage; sharing (CPU) or CUDA-IPC
        pickling then keeps all processes on the same weights and WeightBlob._adopt re-finds the buffer in each child."""
        dev = next(self.parameters()).device
        self.blob.ensure(dev)
        return super().share_memory()

    def __getstate__(self):       # the blob holds ctypes handles; it is rebuilt lazily after unpickling
        state = self.__dict__.copy()
        state['_blob'] = None
        return state

    def forward(self, p, npc, stage, npc_geo_feats, npc_col_feats, pts_num=16, is_tracker=False, cloud_pos=None,
                pts_views_d=None, dynamic_r_query=None, exposure_feat=None):
        """Per-point decode, forward only: raw (P,4) [r,g,b,occ], ray_mask (P/pts_num,), point_mask (P,)."""
        from .renderer import decode_points
        return decode_points(self, p, npc, stage, npc_geo_feats, npc_col_feats, pts_num, cloud_pos,
                             dynamic_r_query, exposure_feat)
