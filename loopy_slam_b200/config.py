"""Mirror of /root/reference/src/config.py:10-57 (YAML with recursive ``inherit_from``) and
/root/reference/src/conv_onet/config.py:4-22 (model factory)."""
import yaml

from .decoder import NICER


def _read_yaml(path):
    with open(path, 'r') as f:
        return yaml.full_load(f) or {}


def deep_merge(dst, src):
    """Merge `src` into `dst` in place, recursing into nested mappings (explicit stack, no recursion)."""
    stack = [(dst, src)]
    while stack:
        d, s_ = stack.pop()
        for key, val in s_.items():
            if isinstance(val, dict):
                node = d.get(key)
                if not isinstance(node, dict):
                    node = d[key] = {}
                stack.append((node, val))
            else:
                d[key] = val
    return dst


def load_config(path, default_path=None):
    """YAML config with `inherit_from` chains, same resolution order as the reference loader
    (/root/reference/src/config.py:10-57): the root of the chain (or `default_path` when the chain ends without one) is
    the base, every file further down the chain overrides it."""
    chain, cur = [], path
    while cur is not None:
        doc = _read_yaml(cur)
        chain.append(doc)
        cur = doc.get('inherit_from')
    cfg = _read_yaml(default_path) if default_path is not None else {}
    for doc in reversed(chain):
        deep_merge(cfg, doc)
    return cfg


def update_recursive(dict1, dict2):   # reference-compatible name (src/config.py:45)
    deep_merge(dict1, dict2)


def get_model(cfg):
    return NICER(cfg=cfg, dim=cfg['data']['dim'], c_dim=cfg['model']['c_dim'],
                 pos_embedding_method=cfg['model']['pos_embedding_method'],
                 use_view_direction=cfg['use_view_direction'])


def default_cfg(dataset='replica'):
    """The hot-path subset of configs/point_slam.yaml merged with the per-dataset overrides
    (configs/Replica/replica.yaml, configs/TUM_RGBD/tum.yaml, configs/ScanNet/scannet.yaml), for
    benches/tests that run where the reference tree is absent."""
    cfg = {
        'use_view_direction': False, 'use_dynamic_radius': True, 'setup_seed': 1219,
        'data': {'dim': 3},
        'model': {'c_dim': 32, 'exposure_dim': 8, 'pos_embedding_method': 'fourier', 'encode_rel_pos_in_col': True,
                  'encode_exposure': False, 'encode_viewd': True},
        'mapping': {'device': 'cuda:0', 'w_color_loss': 0.1, 'pixels': 1000, 'mapping_window_size': 5, 'iters': 400},
        'tracking': {'device': 'cuda:0', 'w_color_loss': 0.5, 'pixels': 200, 'iters': 20, 'lr': 0.002,
                     'ignore_edge_W': 20, 'ignore_edge_H': 20},
        'cam': {'H': 680, 'W': 1200, 'fx': 600.0, 'fy': 600.0, 'cx': 599.5, 'cy': 339.5, 'crop_edge': 0},
        'rendering': {'N_surface': 5, 'near_end': 0.3, 'near_end_surface': 0.98, 'far_end_surface': 1.02,
                      'sigmoid_coef_tracker': 0.1, 'sigmoid_coef_mapper': 0.1, 'sample_near_pcl': True,
                      'skip_zero_depth_pixel': False},
        'pointcloud': {'nn_num': 8, 'min_nn_num': 2, 'N_add': 3, 'nn_weighting': 'distance', 'radius_add': 0.04,
                       'radius_min': 0.02, 'radius_query': 0.08, 'radius_mesh': 0.08, 'radius_add_max': 0.08,
                       'radius_add_min': 0.02, 'radius_query_ratio': 2, 'color_grad_threshold': 0.15,
                       'near_end_surface': 0.98, 'far_end_surface': 1.02, 'nlist': 400, 'nprobe': 4,
                       'fix_interval_when_add_along_ray': False},
    }
    over = {
        'replica': {'use_dynamic_radius': False, 'rendering': {'sample_near_pcl': False},
                    'tracking': {'pixels': 1500, 'iters': 40, 'ignore_edge_W': 100, 'ignore_edge_H': 100},
                    'mapping': {'pixels': 5000, 'mapping_window_size': 12, 'iters': 300}},
        'tum': {'use_dynamic_radius': True, 'model': {'encode_rel_pos_in_col': False},
                'rendering': {'sample_near_pcl': False},
                'cam': {'H': 480, 'W': 640, 'fx': 517.3, 'fy': 516.5, 'cx': 318.6, 'cy': 255.3, 'crop_edge': 8},
                'tracking': {'pixels': 5000, 'iters': 200}, 'mapping': {'pixels': 10000, 'mapping_window_size': 10}},
        'scannet': {'use_dynamic_radius': True,
                    'model': {'encode_rel_pos_in_col': False, 'encode_exposure': True, 'encode_viewd': False},
                    'rendering': {'sample_near_pcl': False, 'near_end_surface': 0.96, 'far_end_surface': 1.04},
                    'pointcloud': {'near_end_surface': 0.96, 'far_end_surface': 1.04},
                    'cam': {'H': 480, 'W': 640, 'fx': 577.6, 'fy': 578.7, 'cx': 318.9, 'cy': 242.7, 'crop_edge': 10},
                    'tracking': {'pixels': 5000, 'iters': 100}, 'mapping': {'pixels': 10000, 'mapping_window_size': 20}},
    }[dataset]
    update_recursive(cfg, over)
    return cfg
