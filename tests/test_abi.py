"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol
include/vo_b200.h declares; without a GPU the product fails loudly instead of falling back."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from ros_stereo_slam_b200 import _lib
    return _lib


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "vo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(vo_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    l = lib.load()
    for name in sorted(declared):
        assert hasattr(l, name), name
    assert declared == set(lib.SYMBOLS)
    assert l.vo_abi_version() == 1


def test_default_params_are_reference_constants(lib):
    p = lib.default_params()
    # include/visualSLAM.h:68,82-87; tracking.cpp:34,75; keyFrameManagement.cpp:84-93; VisualSLAM.cpp:120
    assert (p.fx, p.fy, p.cx, p.cy, p.baseline) == (718.856, 718.856, 607.1928, 185.2157, 0.54)
    assert (p.lk_win, p.lk_max_level, p.lk_max_iters, p.lk_eps, p.lk_min_eig) == (21, 3, 30, 0.01, 1e-4)
    assert (p.grid_step, p.f_thr_stereo, p.f_thr_temporal, p.f_conf, p.f_max_iters) == (30, 3.0, 1.0, 0.99, 1000)
    assert (p.pnp_iters, p.pnp_thr, p.pnp_conf) == (100, 1.0, 0.99)
    assert (p.pnp_retry_iters, p.pnp_retry_thr, p.pnp_retry_conf, p.pnp_min_inliers) == (100, 8.0, 0.98, 10)
    assert p.kf_min_inliers == 200


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ros_stereo_slam_b200 import VisualFrontEnd, VoError
    with pytest.raises(VoError) as e:
        VisualFrontEnd()
    assert e.value.code == lib.VO_ERR_NO_DEVICE


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ros_stereo_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "cv2" not in src or f.endswith((".cu", ".cuh")), f
