"""Locates the reference package (multi_robot_multi_goal_planning) and makes it importable.

The B200 environments are plug-ins of the reference: they derive from its `BaseProblem` and are driven by its
unmodified planners.  Search order:
  1. already importable (a user's own install of the reference);
  2. `baseline/_ref/` at the repository root -- the offline install made by `__graft_entry__.build()` (the tree is
     git-ignored and travels to the GPU box with the snapshot);
  3. `/root/reference/src` (the read-only checkout of the build container).
The reference imports GUI / solver packages at module level that the planners never touch on this path
(matplotlib, simple_parsing, rai's `robotic`; recipe: SURVEY.md 8c): modules that are not installed are replaced
by inert stand-ins, installed ones are left alone.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
from unittest.mock import MagicMock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALL_DIR = os.path.join(ROOT, "baseline", "_ref")
CHECKOUT_SRC = "/root/reference/src"
PKG = "multi_robot_multi_goal_planning"
OPTIONAL = ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.patches", "matplotlib.collections",
            "mpl_toolkits", "mpl_toolkits.mplot3d", "robotic", "simple_parsing")


def _stub_missing() -> None:
    for name in OPTIONAL:
        if name in sys.modules:
            continue
        top = name.split(".")[0]
        try:
            found = importlib.util.find_spec(top) is not None and not isinstance(sys.modules.get(top), MagicMock)
        except (ImportError, ValueError):
            found = False
        if not found:
            sys.modules[name] = MagicMock()


def reference_path():
    """directory to put on sys.path, or None if the reference is importable as is / nowhere to be found"""
    for cand in (INSTALL_DIR, CHECKOUT_SRC):
        if os.path.isdir(os.path.join(cand, PKG)):
            return cand
    return None


def ensure_reference() -> bool:
    """True if `import multi_robot_multi_goal_planning` works afterwards."""
    if PKG in sys.modules:
        return True
    _stub_missing()
    try:
        if importlib.util.find_spec(PKG) is None:
            raise ImportError
    except (ImportError, ValueError):
        p = reference_path()
        if p is None:
            return False
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        importlib.import_module(PKG)
        return True
    except Exception:
        return False


def install_reference(force: bool = False) -> str:
    """The offline install of the reference for the GPU box (`baseline/_ref`, git-ignored, not gpurun-ignored).
    `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference`
    fails here (the build backend `hatchling` is neither installed nor in the wheelhouse), so the package's Python
    files are copied as they are -- the reference is pure Python; its 184 MB of mesh / model assets are not needed
    by the b200 environments and stay behind."""
    import shutil
    src = os.path.join(CHECKOUT_SRC, PKG)
    dst = os.path.join(INSTALL_DIR, PKG)
    if not os.path.isdir(src):
        return dst if os.path.isdir(dst) else ""
    if os.path.isdir(dst) and not force:
        newest_src = max(os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(src) for f in fs if f.endswith(".py"))
        stamp = os.path.join(INSTALL_DIR, ".installed")
        if os.path.exists(stamp) and os.path.getmtime(stamp) >= newest_src:
            return dst
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=lambda d, names: [n for n in names if n == "assets" or n == "__pycache__" or
                                                       (os.path.isfile(os.path.join(d, n)) and not n.endswith(".py"))])
    # the experiment configurations the bench quotes (configs/experiments-ral/box_stacking.json) ride along
    cfg = os.path.join(os.path.dirname(CHECKOUT_SRC), "configs", "experiments-ral")
    if os.path.isdir(cfg):
        shutil.copytree(cfg, os.path.join(INSTALL_DIR, "configs", "experiments-ral"), dirs_exist_ok=True)
    with open(os.path.join(INSTALL_DIR, ".installed"), "w") as f:
        f.write("copied from /root/reference/src (pure-Python package, assets excluded)\n")
    return dst
