"""Frustum feature selection on the device: which neural points may be
optimised for the current frame.  Restates Mapper.get_mask_from_c2w
(/root/reference/src/Mapper.py:165-217: CPU numpy projection of ALL points + cv2.remap per mapped
frame, returning a Python list) as a handful of device tensor ops returning an index tensor."""
import torch
import torch.nn.functional as F


def get_mask_from_c2w(cloud_pos, c2w, depth, H, W, fx, fy, cx, cy, edge=-4):
    """-> int64 indices of cloud rows that project inside the (edge-cropped) image and lie in front of
    the camera no deeper than the bilinearly sampled sensor depth + 0.5 m (zero depth -> max depth)."""
    dev = cloud_pos.device
    c2w = c2w.to(device=dev, dtype=torch.float32)
    if c2w.shape[0] == 3:
        c2w = torch.cat([c2w, torch.tensor([[0., 0., 0., 1.]], device=dev)], 0)
    w2c = torch.linalg.inv(c2w)
    cam = cloud_pos.float() @ w2c[:3, :3].t() + w2c[:3, 3]
    z = cam[:, 2] + 1e-5                      # negative in front of the camera
    u = (fx * (-cam[:, 0]) + cx * cam[:, 2]) / z
    v = (fy * cam[:, 1] + cy * cam[:, 2]) / z
    # bilinear lookup of the sensor depth at (u, v), pixel centres at integer coordinates
    gx = (u / (W - 1)) * 2 - 1
    gy = (v / (H - 1)) * 2 - 1
    grid = torch.stack([gx, gy], -1).reshape(1, 1, -1, 2)
    d = F.grid_sample(depth.float().reshape(1, 1, H, W), grid, mode='bilinear', padding_mode='zeros',
                      align_corners=True).reshape(-1)
    d = torch.where(d == 0, d.max(), d)
    mask = (u < W - edge) & (u > edge) & (v < H - edge) & (v > edge) & (-z >= 0) & (-z <= d + 0.5)
    return torch.nonzero(mask, as_tuple=True)[0]
