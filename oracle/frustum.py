"""ORACLE (test infrastructure, never imported by the product): Mapper.get_mask_from_c2w restated with the reference's own
numpy + cv2 calls (/root/reference/src/Mapper.py:165-217), cv2.remap included (cv2 is available on both boxes).
PINNED: tests/golden/make_golden_frustum.py runs the REAL Mapper.get_mask_from_c2w / filter_point_before_add (unbound, on a
stand-in object) and tests/test_oracle_golden.py checks this restatement against those row ids bit for bit."""
import numpy as np


def get_mask_from_c2w(points, c2w, depth_np, H, W, fx, fy, cx, cy, edge):
    import cv2
    points = np.array(points, dtype=np.float64).reshape(-1, 3)                     # :178 (python floats -> float64)
    c2w = np.asarray(c2w, dtype=np.float32)
    if c2w.shape[0] == 3:
        c2w = np.concatenate([c2w, np.array([[0, 0, 0, 1]], np.float32)], 0)
    w2c = np.linalg.inv(c2w)                                                       # :181
    ones = np.ones_like(points[:, 0]).reshape(-1, 1)
    homo_vertices = np.concatenate([points, ones], axis=1).reshape(-1, 4, 1)
    cam_cord_homo = w2c @ homo_vertices
    cam_cord = cam_cord_homo[:, :3]
    K = np.array([[fx, .0, cx], [.0, fy, cy], [.0, .0, 1.0]]).reshape(3, 3)
    cam_cord[:, 0] *= -1                                                           # :189
    uv = K @ cam_cord
    z = uv[:, -1:] + 1e-5
    uv = uv[:, :2] / z
    uv = uv.astype(np.float32)
    depths = []
    chunk = int(3e4)
    depth_np = np.asarray(depth_np, dtype=np.float32)
    for i in range(0, uv.shape[0], chunk):                                         # :196-202
        depths += [cv2.remap(depth_np, uv[i:i + chunk, 0], uv[i:i + chunk, 1], interpolation=cv2.INTER_LINEAR)[:, 0].reshape(-1, 1)]
    depths = np.concatenate(depths, axis=0)
    mask = (uv[:, 0] < W - edge) * (uv[:, 0] > edge) * (uv[:, 1] < H - edge) * (uv[:, 1] > edge)
    zero_mask = (depths == 0)
    depths[zero_mask] = np.max(depths)
    mask = mask & (0 <= -z[:, :, 0]) & (-z[:, :, 0] <= depths + 0.5)
    return np.where(mask.reshape(-1))[0]
