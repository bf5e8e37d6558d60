"""ros_stereo_slam_b200 -- B200-native visual-odometry front-end (hand-written CUDA, sm_100a)
behind the C ABI of include/vo_b200.h.  Drop-in for the per-frame hot path of
Gautham-JS/ROS_Stereo_SLAM: grid keypoints (+ANMS), pyramidal LK with F-matrix RANSAC
rejection, stereo correspondence + DLT triangulation, PnP-RANSAC pose.

The package holds only what that path needs: csrc/ (kernels + C ABI), build.py, and the
host-side mirror of the reference interface (frontend.VisualFrontEnd).  It never imports
oracle/ and has no CPU compute path."""
from . import _lib  # noqa: F401
from .frontend import VisualFrontEnd, VoError  # noqa: F401

__all__ = ["VisualFrontEnd", "VoError"]
