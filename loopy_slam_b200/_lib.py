"""ctypes binding of the C ABI in include/lsr.h (built in-tree as loopy_slam_b200/liblsr.so).

There is deliberately NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised -- the product path never routes through the oracle or PyTorch eager.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('LSR_LIB', os.path.join(_HERE, 'liblsr.so'))   # LSR_LIB: A/B experiments only

LSR_STAGE = {'geometry': 0, 'color': 1}
FLAG_REL_POS, FLAG_DYNAMIC_R, FLAG_SKIP_ZERO_DEPTH, FLAG_SAMPLE_NEAR_PCL, FLAG_SAVE_LIGHT, FLAG_FWD_ONLY = 1, 2, 4, 8, 16, 32
RGB_SIGMOID, RGB_RAW, RGB_AFFINE_SIGMOID = 0, 1, 2
GRAD_GEO_FEATS, GRAD_COL_FEATS, GRAD_GEO_W, GRAD_GEO_B, GRAD_COL_W, GRAD_RAYS, GRAD_AFFINE = 1, 2, 4, 8, 16, 32, 64

# every exported symbol declared in include/lsr.h (checked by tests/test_abi.py)
EXPORTS = [
    'lsr_version', 'lsr_strerror', 'lsr_device_sm_count', 'lsr_launch_count', 'lsr_grid_workspace_bytes', 'lsr_grid_build',
    'lsr_knn_query', 'lsr_sample_rays', 'lsr_sample_rays_filtered', 'lsr_sample_rays_filtered_sync', 'lsr_sample_rays_bwd', 'lsr_pose_fwd', 'lsr_pose_bwd',
    'lsr_render_workspace_bytes', 'lsr_far_bound', 'lsr_render_fwd', 'lsr_render_bwd', 'lsr_dynamic_radius', 'lsr_frustum_scratch_bytes', 'lsr_frustum_mask',
    'lsr_loss_scratch_bytes', 'lsr_mapper_loss', 'lsr_tracker_resid', 'lsr_tracker_loss', 'lsr_debug_program_stats',
]


class LsrParams(ctypes.Structure):
    _fields_ = [('n_surface', ctypes.c_int32), ('nn_num', ctypes.c_int32), ('min_nn_num', ctypes.c_int32),
                ('c_dim', ctypes.c_int32), ('near_end_surface', ctypes.c_float),
                ('far_end_surface', ctypes.c_float), ('near_end', ctypes.c_float),
                ('sigmoid_coef', ctypes.c_float), ('radius_query', ctypes.c_double),
                ('flags', ctypes.c_int32), ('rgb_mode', ctypes.c_int32)]


class LsrWeights(ctypes.Structure):
    _fields_ = [('blob', ctypes.c_void_p), ('n_elems', ctypes.c_int64),
                ('g_fc_w', ctypes.c_int32 * 5), ('g_fc_b', ctypes.c_int32 * 5), ('g_B', ctypes.c_int32),
                ('g_lin_w', ctypes.c_int32 * 5), ('g_lin_b', ctypes.c_int32 * 5),
                ('g_out_w', ctypes.c_int32), ('g_out_b', ctypes.c_int32),
                ('c_fc_w', ctypes.c_int32 * 5), ('c_fc_b', ctypes.c_int32 * 5), ('c_B', ctypes.c_int32),
                ('c_Brel', ctypes.c_int32), ('c_nb1_w', ctypes.c_int32), ('c_nb1_b', ctypes.c_int32),
                ('c_nb2_w', ctypes.c_int32), ('c_nb2_b', ctypes.c_int32),
                ('c_lin_w', ctypes.c_int32 * 5), ('c_lin_b', ctypes.c_int32 * 5),
                ('c_out_w', ctypes.c_int32), ('c_out_b', ctypes.c_int32)]


_lib = None


def lib():
    """The loaded shared library (raises if it was not built: run `python -c "import
    __graft_entry__ as g; g.build()"` first)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} not found: the CUDA extension is not built '
                               '(run __graft_entry__.build()); there is no CPU fallback')
        L = ctypes.CDLL(LIB_PATH)
        L.lsr_strerror.restype = ctypes.c_char_p
        L.lsr_strerror.argtypes = [ctypes.c_int]
        vp, i64, i32, f32, f64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_double
        L.lsr_device_sm_count.argtypes = [ctypes.POINTER(ctypes.c_int)]
        L.lsr_launch_count.argtypes = [ctypes.c_int]
        L.lsr_grid_workspace_bytes.argtypes = [i64, i64, ctypes.POINTER(ctypes.c_size_t)]
        L.lsr_grid_build.argtypes = [vp, i64, f32, i64, vp, ctypes.c_size_t, vp]
        L.lsr_knn_query.argtypes = [vp, vp, vp, f64, i64, vp, vp, vp, vp]
        L.lsr_sample_rays.argtypes = [vp, vp, i32, i32, f32, f32, f32, f32, vp, i32, vp, i64, i32, i32, i32, i32,
                                      vp, vp, vp, vp, vp, vp, vp]
        L.lsr_sample_rays_filtered.argtypes = [vp, vp, i32, i32, f32, f32, f32, f32, vp, i32, vp, i64, i32, i32, i32, i32, f32,
                                               vp, vp, vp, vp, vp, vp, vp, vp]
        L.lsr_sample_rays_filtered_sync.argtypes = [vp, vp, i32, i32, f32, f32, f32, f32, vp, i32, vp, i64, i32, i32, i32, i32, f32,
                                                    vp, vp, vp, vp, vp, vp, ctypes.POINTER(ctypes.c_int32), vp]
        L.lsr_sample_rays_bwd.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, vp, vp]
        L.lsr_pose_fwd.argtypes = [vp, vp, vp]
        L.lsr_pose_bwd.argtypes = [vp, vp, vp, vp]
        L.lsr_render_workspace_bytes.argtypes = [ctypes.POINTER(LsrParams), i64, ctypes.c_int,
                                                 ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
        L.lsr_far_bound.argtypes = [vp, i64, i64, vp, vp]
        L.lsr_dynamic_radius.argtypes = [vp, vp, i32, i32, f64, f64, f64, f64, vp, vp, vp]
        L.lsr_frustum_scratch_bytes.argtypes = [i64, ctypes.POINTER(ctypes.c_size_t)]
        L.lsr_frustum_mask.argtypes = [vp, i64, ctypes.POINTER(ctypes.c_double), vp, i32, i32, f64, f64, f64, f64, i32, vp, vp, vp]
        L.lsr_render_fwd.argtypes = [ctypes.POINTER(LsrParams), vp, vp, i64, vp, vp, vp, vp, vp, i64, vp, i64, vp, vp,
                                     vp, vp, vp,
                                     ctypes.POINTER(LsrWeights), vp, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp]
        L.lsr_render_bwd.argtypes = [ctypes.POINTER(LsrParams), vp, vp, i64, vp, vp, vp, vp, i64, vp, vp,
                                     vp, vp, vp,
                                     ctypes.POINTER(LsrWeights), vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp,
                                     ctypes.c_int, vp, vp, vp, vp, vp, vp, vp]
        L.lsr_loss_scratch_bytes.argtypes = [ctypes.POINTER(ctypes.c_size_t)]
        L.lsr_mapper_loss.argtypes = [vp, vp, vp, vp, vp, i64, ctypes.c_int, f32, vp, vp, vp, vp, vp]
        L.lsr_tracker_resid.argtypes = [vp, vp, vp, i64, ctypes.c_int, vp, vp, vp]
        L.lsr_tracker_loss.argtypes = [vp, vp, vp, vp, vp, vp, i64, vp, ctypes.c_int, f32, vp, vp, vp, vp, vp, vp]
        for name in EXPORTS:
            if name not in ('lsr_strerror', 'lsr_launch_count'):
                getattr(L, name).restype = ctypes.c_int
        L.lsr_launch_count.restype = ctypes.c_longlong
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f'{what} failed: {lib().lsr_strerror(rc).decode()} (code {rc})')


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)
_get_device = getattr(torch._C, '_cuda_getDevice', None)


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on `device` (the raw-handle fast path avoids building a Stream object:
    this runs ~15 times per mapping iteration)."""
    if _raw_stream is not None and isinstance(device, torch.device) and device.index is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def on_device(device):
    """`with on_device(dev):` == `with torch.cuda.device(dev):`, skipping the device switch when dev already is current."""
    if _get_device is not None and isinstance(device, torch.device) and device.index is not None and device.index == _get_device():
        return _NO_GUARD
    return torch.cuda.device(device)


def require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor: the lsr renderer has no CPU path')
