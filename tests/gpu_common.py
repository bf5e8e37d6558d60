import os
import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vo_golden_v1.npz")


def golden():
    return np.load(GOLDEN)


def make_frontend(**kw):
    from ros_stereo_slam_b200 import VisualFrontEnd
    return VisualFrontEnd(**kw)
