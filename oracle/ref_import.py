"""Import the *real* Loopy-SLAM reference as a library (build container only).

TEST INFRASTRUCTURE.  Works only where ``/root/reference`` exists (never on the GPU box);
used by ``tests/golden/make_golden.py`` to mint golden vectors and by the optional
cross-check test ``tests/test_oracle_vs_reference.py``.

Recipe: SURVEY.md Appendix C.  Several third-party imports of the reference are absent
here (open3d, faiss, skimage, pydbow3, turtle/tkinter, ...) but are not needed by the
hot path, so they are replaced by permissive dummy modules before importing
``src.common``, ``src.conv_onet.models.decoder``, ``src.utils.Renderer`` and ``src.config``.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LOOPY_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "turtle", "open3d", "open3d.core", "skimage", "skimage.color", "skimage.filters",
    "faiss", "faiss.contrib", "faiss.contrib.torch_utils", "pydbow3", "matplotlib",
    "matplotlib.pyplot", "colorama", "torchmetrics", "torchmetrics.image",
    "torchmetrics.image.lpip", "pytorch_msssim",
]


class _Permissive(types.ModuleType):
    """A module whose every attribute is another permissive module / no-op callable."""

    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        return _Permissive(self.__name__ + "." + key)

    def __call__(self, *a, **k):
        return None


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def import_reference():
    """Returns (common, decoder, Renderer_mod, config) modules of the real reference."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = _Permissive(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import src.common as common  # noqa
    import src.conv_onet.models.decoder as decoder  # noqa
    import src.utils.Renderer as renderer  # noqa
    import src.config as config  # noqa
    return common, decoder, renderer, config


def load_cfg(scene_yaml="configs/Replica/room0.yaml"):
    """Load a reference YAML (inherit_from paths are relative to the reference root)."""
    _, _, _, config = import_reference()
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        cfg = config.load_config(scene_yaml, "configs/point_slam.yaml")
    finally:
        os.chdir(cwd)
    return cfg


def build_model(cfg, load_pretrained=True):
    """NICER from the reference factory; geo decoder from pretrained/middle_fine.pt
    (mirrors src/Point_SLAM.py:185-198)."""
    import torch
    _, _, _, config = import_reference()
    model = config.get_model(cfg)
    if load_pretrained:
        ckpt = torch.load(os.path.join(REFERENCE_ROOT, "pretrained/middle_fine.pt"),
                          map_location="cpu", weights_only=False)
        geo = {}
        for key, val in ckpt["model"].items():
            if key.startswith("coarse.decoder."):
                geo[key[len("coarse.decoder."):]] = val
        model.geo_decoder.load_state_dict(geo, strict=False)
    return model
