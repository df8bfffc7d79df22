"""TEST INFRASTRUCTURE ONLY -- import the reference's own functions in the build container.

``/root/reference/MonocularSlam/slam.py`` is pure Python but imports ``pangolin``,
``OpenGL.GL`` and ``skimage`` at module top (slam.py:15,17,20,21) and uses ``np.mat``
(slam.py:168), removed in NumPy 2.  Stubbing those lets ``normalize``, ``match_points``,
``extract_Rt``, ``triangulate`` and ``check_points`` run as the REAL reference code.
``/root/reference`` does not exist on the GPU box, so this loader is used only by
``tests/golden/make_golden.py`` and by CPU tests that skip when the tree is absent.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REF_SLAM = "/root/reference/MonocularSlam/slam.py"


def available():
    return os.path.isfile(REF_SLAM)


def load():
    """Return the reference module (functions only; its __main__ block does not run)."""
    if not available():
        raise FileNotFoundError(REF_SLAM)
    for name in ("pangolin", "OpenGL", "OpenGL.GL", "skimage", "skimage.measure",
                 "skimage.transform"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["skimage.measure"].ransac = None
    sys.modules["skimage.transform"].EssentialMatrixTransform = None
    sys.modules["OpenGL"].GL = sys.modules["OpenGL.GL"]
    if not hasattr(np, "mat"):
        np.mat = np.asmatrix
    spec = importlib.util.spec_from_file_location("openwork_reference_slam", REF_SLAM)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
