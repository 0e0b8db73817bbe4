"""loopy_slam_b200 -- B200-native (sm_100a) neural-point volume renderer for Loopy-SLAM's
per-iteration hot path, behind the reference's own Python surface:

    Renderer.render_batch_ray / render_img / eval_points      (renderer.py)
    NICER decoder parameter container                         (decoder.py)
    NeuralPointCloud.find_neighbors_faiss / add_neural_points (neural_point.py)
    get_samples / get_rays / get_camera_from_tensor           (common.py)

All arithmetic runs in hand-written CUDA (csrc/) behind the C ABI of include/lsr.h.
"""
from . import _lib  # noqa: F401
from .config import get_model, load_config, default_cfg  # noqa: F401
from .decoder import NICER  # noqa: F401
from .renderer import Renderer, GridIndex, FeatureSubset  # noqa: F401
from .neural_point import NeuralPointCloud  # noqa: F401
from .loss import mapper_loss, tracker_loss  # noqa: F401
from .common import (get_samples, get_rays, get_rays_from_uv, get_camera_from_tensor, quad2rotation,  # noqa: F401
                     raw2outputs_nerf_color)

__version__ = '0.1.0'
