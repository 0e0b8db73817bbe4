"""Mint the golden vectors of the frustum feature selection by running the REAL reference methods
(/root/reference/src/Mapper.py: get_mask_from_c2w :165-217, filter_point_before_add :137-163) as unbound functions on a
stand-in object (numpy + cv2.remap, as in the reference).  Build-container only; writes tests/golden/frustum.npz.

    python tests/golden/make_golden_frustum.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'frustum.npz')


def main():
    ref_import.import_reference()
    import src.Mapper as M
    from loopy_slam_b200.stream import SyntheticRoom, build_point_cloud, sample_batch
    room = SyntheticRoom(H=96, W=128, fx=80., fy=80., cx=63.5, cy=47.5, n_frames=100, half=(1.0, 0.8, 0.6), hole_frac=0.02)
    cloud, _, _ = build_point_cloud(room, 6000, pixels_per_frame=3000, frame_ids=[0, 10, 20, 30, 50, 70], max_frames=60)
    s = types.SimpleNamespace(H=room.H, W=room.W, fx=room.fx, fy=room.fy, cx=room.cx, cy=room.cy, device='cpu')
    s.npc = types.SimpleNamespace(get_cloud_pos=lambda: cloud.tolist())
    out = dict(intr=np.array([room.H, room.W, room.fx, room.fy, room.cx, room.cy]), cloud=cloud.numpy())
    cases = [(5, -4), (20, 0), (75, 10)]
    out['cases'] = np.array(cases)
    for k, (fid, edge) in enumerate(cases):
        color, depth, c2w = room.frame(fid)
        s.frustum_edge = edge
        idx = M.Mapper.get_mask_from_c2w(s, c2w, depth.numpy())
        out[f'c2w{k}'], out[f'depth{k}'], out[f'idx{k}'] = c2w.numpy(), depth.numpy(), np.array(idx, dtype=np.int64)
        print('frame', fid, 'edge', edge, 'selected', len(idx), 'of', cloud.shape[0])
    # filter_point_before_add: samples of frame 30 seen from the camera of frame 24
    o, d, g, _ = sample_batch(room, [30], 800, seed=5)
    prev = room.frame(24)[2]
    m = M.Mapper.filter_point_before_add(s, o, d, g, prev)
    out.update(f_o=o.numpy(), f_d=d.numpy(), f_g=g.numpy(), f_prev=prev.numpy(), f_mask=m.numpy())
    print('filter_point_before_add: outside', int(m.sum()), 'of', m.numel())
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
