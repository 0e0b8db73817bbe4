"""CPU: the C-ABI shared library loads and exports every symbol include/lsr.h declares; argument
validation works without a GPU (no compute call is issued)."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    entry.build()
    from loopy_slam_b200 import _lib
    return _lib.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, 'include', 'lsr.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(lsr_[a-z_0-9]+)\s*\(', hdr))
    assert len(declared) >= 14
    from loopy_slam_b200 import _lib
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layout_matches_header():
    from loopy_slam_b200 import _lib
    assert ctypes.sizeof(_lib.LsrParams) == 48
    assert ctypes.sizeof(_lib.LsrWeights) == 16 + 4 * (10 + 1 + 10 + 2 + 10 + 2 + 4 + 10 + 2) + 4   # 8-byte tail pad
    assert _lib.LsrParams.radius_query.offset == 32


def test_argument_validation_without_gpu(lib):
    from loopy_slam_b200 import _lib
    out = ctypes.c_size_t()
    assert lib.lsr_grid_workspace_bytes(1000, 1 << 16, ctypes.byref(out)) == 0 and out.value > 1000 * 16
    assert lib.lsr_grid_workspace_bytes(-1, 1 << 16, ctypes.byref(out)) == 1
    prm = _lib.LsrParams(n_surface=5, nn_num=8, min_nn_num=2, c_dim=32, flags=1)
    sb, cb = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.lsr_render_workspace_bytes(ctypes.byref(prm), 4992, 1, ctypes.byref(sb), ctypes.byref(cb)) == 0
    assert sb.value > 4992 * 5 * 4 * 2500 and cb.value > 400000
    full_scratch, full_saved = cb.value, sb.value
    assert full_saved < 4992 * 5 * 4 * 2900                     # the h planes are not saved: 11.3 KB per sample
    prm.flags = 1 | _lib.FLAG_FWD_ONLY                          # forward-only: scratch without the backward's hand-over planes
    assert lib.lsr_render_workspace_bytes(ctypes.byref(prm), 4992, 1, ctypes.byref(sb), ctypes.byref(cb)) == 0
    assert cb.value < full_scratch / 4 and cb.value > 4992 * 5 * 120
    prm.flags = 1 | _lib.FLAG_SAVE_LIGHT
    assert lib.lsr_render_workspace_bytes(ctypes.byref(prm), 4992, 1, ctypes.byref(sb), ctypes.byref(cb)) == 0
    assert sb.value < 4992 * 5 * 200                            # light save: 148 B per sample (+ padding)
    prm.flags = 1
    prm.nn_num = 7
    assert lib.lsr_render_workspace_bytes(ctypes.byref(prm), 10, 1, ctypes.byref(sb), ctypes.byref(cb)) == 4
    assert lib.lsr_strerror(4) == b'unsupported configuration'
    assert lib.lsr_pose_fwd(None, None, None) == 1
    assert lib.lsr_loss_scratch_bytes(ctypes.byref(out)) == 0 and out.value >= 32
    assert lib.lsr_mapper_loss(None, None, None, None, None, 10, 1, 0.1, None, None, None, None, None) == 1
    assert lib.lsr_tracker_loss(None, None, None, None, None, None, 10, None, 1, 0.5, None, None, None, None, None, None) == 1


def test_no_cpu_fallback():
    import torch
    import loopy_slam_b200 as L
    cfg = L.default_cfg('replica')
    from parity import SlamLike
    r = L.Renderer(cfg, None, SlamLike(64, 64, 40, 40, 31.5, 31.5))
    m = L.get_model(cfg)
    with pytest.raises(RuntimeError, match='no CPU path'):
        r.render_batch_ray(None, m, torch.zeros(4, 3), torch.zeros(4, 3), 'cpu', 'color', gt_depth=torch.ones(4),
                           npc_geo_feats=torch.zeros(8, 32), npc_col_feats=torch.zeros(8, 32),
                           cloud_pos=torch.zeros(8, 3))


def test_library_sass_is_tcgen05_only():
    """Every MLP layer of both passes runs on the 5th-generation tensor cores: the sm_100a SASS of the library holds
    tcgen05 MMAs (UTCHMMA) with TMEM loads / stores (LDTM / STTM) and bulk copies (UBLKCP), and no legacy mma.sync (HMMA)."""
    import re
    import shutil
    import subprocess
    from loopy_slam_b200 import _lib
    exe = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(exe):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([exe, '-sass', _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    ops = re.findall(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', sass, flags=re.M)
    names = {o.split('.')[0] for o in ops}
    assert 'UTCHMMA' in names and 'LDTM' in names and 'STTM' in names and 'UBLKCP' in names, sorted(names)[:50]
    assert 'HMMA' not in names and 'IMMA' not in names
    per_kernel = re.split(r'Function : ', sass)[1:]
    with_umma = [k.split('\n')[0] for k in per_kernel if 'UTCHMMA' in k]
    assert any('render_fwd_kernel' in k for k in with_umma) and any('trunk_bwd_umma_kernel' in k for k in with_umma) \
        and any('geo_bwd_umma_kernel' in k for k in with_umma), with_umma
