"""Mint the golden vectors of the pixel/ray sampling functions by running the REAL reference
(/root/reference/src/common.py: get_samples :237-259, get_camera_from_tensor :327-343) on CPU with a
seeded global torch generator.  Build-container only; writes tests/golden/sampling.npz.

    python tests/golden/make_golden_sampling.py
(quad2rotation :314 calls `.to(quad.get_device())`, which fails on CPU tensors: the one line is
re-issued with device=quad.device, as SURVEY.md 8c prescribes.)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'sampling.npz')


def main():
    common, _, _, _ = ref_import.import_reference()
    _gd = torch.Tensor.get_device
    g = torch.Generator().manual_seed(1219)
    H, W = 48, 64
    fx, fy, cx, cy = 55.0, 56.0, 31.5, 23.5
    depth = torch.rand(H, W, generator=g) * 3 + 0.5
    depth[torch.rand(H, W, generator=g) < 0.1] = 0.0           # holes
    color = torch.rand(H, W, 3, generator=g)
    cam = torch.tensor([0.9, 0.1, -0.2, 0.3, 0.5, -0.4, 1.2])   # unnormalised quaternion (w,x,y,z) | T
    torch.Tensor.get_device = lambda t: 'cpu'                   # common.py:314 on CPU tensors
    try:
        c2w = common.get_camera_from_tensor(cam)
    finally:
        torch.Tensor.get_device = _gd
    out = dict(depth=depth.numpy(), color=color.numpy(), cam=cam.numpy(), c2w=c2w.numpy(),
               intr=np.array([H, W, fx, fy, cx, cy], dtype=np.float64))
    cases = [(0, H, 0, W, 300, True, None), (5, H - 7, 9, W - 3, 200, False, None), (4, 40, 4, 60, 256, True, 2.5)]
    for k, (H0, H1, W0, W1, n, filt, lim) in enumerate(cases):
        torch.manual_seed(100 + k)
        # the picks select_uv draws from the global generator (same call, same order)
        idx = torch.randint((H1 - H0) * (W1 - W0), (n,))
        torch.manual_seed(100 + k)
        o, d, sd, sc, i, j = common.get_samples(H0, H1, W0, W1, n, H, W, fx, fy, cx, cy, c2w, depth, color, 'cpu',
                                                depth_filter=filt, return_index=True, depth_limit=lim)
        out[f'case{k}'] = np.array([H0, H1, W0, W1, n, int(filt), -1.0 if lim is None else lim], dtype=np.float64)
        out[f'idx{k}'] = idx.numpy()
        for name, t in (('o', o), ('d', d), ('sd', sd), ('sc', sc), ('i', i), ('j', j)):
            out[f'{name}{k}'] = t.numpy()
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
