"""Shared test helpers: golden loading, error metrics."""
import json
import os

import numpy as np
import torch

from oracle.render import OracleCfg

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

GOLDEN_CASES = [
    'replica_color_mapper', 'replica_geometry_mapper', 'replica_color_tracker',
    'tum_color_mapper_dynr', 'tum_color_tracker_dynr', 'scannet_color_tracker_exposure',
    'scannet_color_mapper_presigmoid', 'replica_color_sparse_zero_depth',
]


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
        self.name = name
        self.raw = {k: z[k] for k in z.files}
        self.stage = str(z['stage'])
        self.is_tracker = bool(int(z['is_tracker']))
        self.ocfg = OracleCfg(**json.loads(str(z['ocfg'])))
        wz = np.load(os.path.join(GOLDEN_DIR, str(z['weights_file'])), allow_pickle=False)
        self.weights = {k: torch.from_numpy(wz[k]) for k in wz.files}
        self.param_grads = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('gw/')}

    def t(self, key, dtype=None):
        v = torch.from_numpy(self.raw[key])
        return v if dtype is None else v.to(dtype)

    def has(self, key):
        return key in self.raw


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)
