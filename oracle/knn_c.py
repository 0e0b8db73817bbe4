"""ctypes wrapper of oracle/c/knn_grid.c (TEST INFRASTRUCTURE; built by __graft_entry__.build()).
Fast exact radius-limited 8-NN on the host CPU -- the neighbour search of the CPU baseline."""
import ctypes
import os
import subprocess

import numpy as np
import torch

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'c')
_SO = os.path.join(_DIR, 'liboracle_knn.so')
_lib = None


def build(force=False):
    src = os.path.join(_DIR, 'knn_grid.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fopenmp', '-shared', '-fPIC', src,
                               '-o', _SO, '-lm'])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        lib.oracle_grid_build.restype = ctypes.c_void_p
        lib.oracle_grid_build.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_float]
        lib.oracle_grid_free.argtypes = [ctypes.c_void_p]
        lib.oracle_knn_query.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double,
                                         ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib = lib
    return _lib


class GridKNN:
    def __init__(self, cloud, cell=0.08):
        self.lib = _load()
        self.cloud = np.ascontiguousarray(cloud.detach().cpu().numpy().astype(np.float32).reshape(-1, 3))
        self.h = self.lib.oracle_grid_build(self.cloud.ctypes.data, self.cloud.shape[0], ctypes.c_float(cell))

    def __del__(self):
        if getattr(self, 'h', None):
            self.lib.oracle_grid_free(self.h)
            self.h = None

    def query(self, q, radius, dynamic_radius=None):
        qn = np.ascontiguousarray(q.detach().cpu().numpy().astype(np.float32).reshape(-1, 3))
        P = qn.shape[0]
        D = np.empty((P, 8), np.float32)
        I = np.empty((P, 8), np.int64)
        n = np.empty((P,), np.int32)
        rd = None
        if dynamic_radius is not None:
            rd = np.ascontiguousarray(dynamic_radius.detach().cpu().numpy().astype(np.float64).reshape(-1))
        self.lib.oracle_knn_query(self.h, qn.ctypes.data, rd.ctypes.data if rd is not None else None,
                                  float(radius), P, D.ctypes.data, I.ctypes.data, n.ctypes.data)
        return torch.from_numpy(D), torch.from_numpy(I), torch.from_numpy(n)
