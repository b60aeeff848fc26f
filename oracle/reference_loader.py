"""ORACLE — test infrastructure only.  Loads the UNMODIFIED reference for timing and fixture generation.

Two places can hold the reference's files:
  * /root/reference                    — the build container (oracle/make_golden.py);
  * <repo>/baseline/_ref/              — a git-ignored copy of the handful of files on the decode path
                                          (BASELINE.md §3), made by `install()` at build time so that
                                          `bench.py --impl reference` can time the real reference on the GPU
                                          box's host cores.  Nothing under baseline/_ref is tracked or edited.
Only bench.py's reference / cpu_baseline legs, tests/ and oracle/make_golden.py import this module.

The three compatibility shims do not touch the decode arithmetic: webcolors / skimage
stubs (imported by utils/utils.py:12 and utils/image.py:18, never called on this path) and a uint8 -> bool cast
for Tensor.masked_select (utils/decode.py:313 passes a uint8 mask, legal in the torch 1.4 the reference pins).
"""
from __future__ import annotations

import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_COPY = os.path.join(ROOT, "baseline", "_ref")
FILES = ["utils/__init__.py", "utils/decode.py", "utils/kmeans.py", "utils/nms.py", "utils/utils.py", "utils/image.py",
         "utils/tranform.py", "utils/cv2_aug_transforms.py", "utils/parell_util.py", "utils/visualize.py", "utils/logger.py",
         "configs/__init__.py", "configs/decode_cfg.yaml", "configs/trans_cfg.json"]
DIRS = ["utils/sync_batchnorm"]


def install(src: str = "/root/reference", dst: str = REF_COPY) -> bool:
    """Copy the decode-path files of the reference into the git-ignored baseline/_ref (build container only)."""
    if not os.path.isdir(src):
        return os.path.isdir(dst)
    for rel in FILES:
        os.makedirs(os.path.dirname(os.path.join(dst, rel)), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), os.path.join(dst, rel))
    for rel in DIRS:
        shutil.copytree(os.path.join(src, rel), os.path.join(dst, rel), dirs_exist_ok=True)
    return True


def available(path: str = REF_COPY) -> bool:
    return all(os.path.exists(os.path.join(path, rel)) for rel in FILES)


def install_shims() -> None:
    import torch
    if "webcolors" not in sys.modules:
        wc = types.ModuleType("webcolors")

        class _RGB:
            red = green = blue = 0
        wc.name_to_rgb = lambda name: _RGB()
        sys.modules["webcolors"] = wc
    if "skimage" not in sys.modules:
        sk, skm = types.ModuleType("skimage"), types.ModuleType("skimage.measure")
        skm.find_contours = lambda *a, **k: []
        sk.measure = skm
        sys.modules["skimage"], sys.modules["skimage.measure"] = sk, skm
    if not getattr(torch.Tensor.masked_select, "_isg_shim", False):
        _ms = torch.Tensor.masked_select

        def masked_select(self, m):
            return _ms(self, m.bool() if m.dtype == torch.uint8 else m)
        masked_select._isg_shim = True
        torch.Tensor.masked_select = masked_select


class Reference:
    """The reference's modules and its stock validation objects (configs/decode_cfg.yaml with draw_flag off,
    configs/trans_cfg.json val transform), imported from `path`."""

    def __init__(self, path: str = REF_COPY):
        if not available(path):
            raise FileNotFoundError("no copy of the reference under %s (run __graft_entry__.build() in the build container)" % path)
        install_shims()
        if path not in sys.path:
            sys.path.insert(0, path)
        for name in [m for m in sys.modules if m == "utils" or m.startswith("utils.") or m == "configs"]:
            mod = sys.modules[name]
            if not getattr(mod, "__file__", "").startswith(path):
                raise RuntimeError("a different top-level `%s` package is already imported (%s)" % (name, getattr(mod, "__file__", "?")))
        from configs import Config, Configer
        from utils import decode, image, kmeans, nms
        from utils.tranform import CommonTransforms, TransInfo
        self.path = path
        self.decode, self.image, self.kmeans, self.nms = decode, image, kmeans, nms
        self.TransInfo = TransInfo
        self.cfg = Config(os.path.join(path, "configs", "decode_cfg.yaml"))
        self.cfg.draw_flag = False
        self.transforms = CommonTransforms(Configer(configs=os.path.join(path, "configs", "trans_cfg.json")), "val")
