"""ctypes binding of libvo_b200.so (C ABI: include/vo_b200.h).

Loading fails loudly when the library has not been built; creating a context fails
loudly (VO_ERR_NO_DEVICE) without a CUDA device.  There is no CPU fallback.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, os.environ.get("VO_B200_LIB", "libvo_b200.so"))

VO_OK = 0
VO_ERR_INVALID_ARG = -1
VO_ERR_NO_DEVICE = -2
VO_ERR_CUDA = -3
VO_ERR_CAPACITY = -4
VO_ERR_TOO_FEW_POINTS = -5
VO_ERR_NO_MODEL = -6
VO_ERR_LOW_INLIERS = -7
VO_ERR_NOT_IMPLEMENTED = -8
VO_ERR_SELF_CHECK = -9

VO_PNP_EPNP5 = 0
VO_PNP_P3P4 = 1

KERNELS = ["pyramid", "lk", "compact", "fmat_solve", "fmat_score", "triangulate", "pnp_solve", "pnp_score",
           "pnp_refine", "select", "misc"]


class VoParams(C.Structure):
    _fields_ = [
        ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
        ("baseline", C.c_double),
        ("width", C.c_int), ("height", C.c_int), ("channels", C.c_int),
        ("lk_win", C.c_int), ("lk_max_level", C.c_int), ("lk_max_iters", C.c_int),
        ("lk_eps", C.c_double), ("lk_min_eig", C.c_double),
        ("grid_step", C.c_int),
        ("f_thr_stereo", C.c_double), ("f_thr_temporal", C.c_double), ("f_conf", C.c_double),
        ("f_max_iters", C.c_int),
        ("pnp_iters", C.c_int), ("pnp_thr", C.c_double), ("pnp_conf", C.c_double),
        ("pnp_retry_iters", C.c_int), ("pnp_retry_thr", C.c_double), ("pnp_retry_conf", C.c_double),
        ("pnp_min_inliers", C.c_int), ("kf_min_inliers", C.c_int), ("ransac_exhaustive", C.c_int), ("f_exhaustive", C.c_int),
        ("max_points", C.c_int), ("max_hypotheses", C.c_int), ("device", C.c_int),
    ]


class VoSgbmParams(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("min_disparity", "num_disparities", "block_size", "p1", "p2", "disp12_max_diff",
                                       "pre_filter_cap", "uniqueness_ratio", "speckle_window_size", "speckle_range")]


class VoFrameResult(C.Structure):
    _fields_ = [
        ("rvec", C.c_double * 3), ("tvec", C.c_double * 3), ("pose3x4", C.c_double * 12),
        ("n_lk_in", C.c_int), ("n_tracked", C.c_int), ("n_inliers", C.c_int), ("attempt_used", C.c_int),
        ("keyframe", C.c_int), ("n_kf_points", C.c_int), ("n_lk_in_stereo", C.c_int),
    ]


class VoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libvo_b200: %s (%d): %s" % (_strerror(code), code, msg))
        self.code = code


_lib = None

# every symbol include/vo_b200.h declares (tests/test_abi.py checks the list against the header)
SYMBOLS = [
    "vo_default_params", "vo_abi_version", "vo_last_error", "vo_strerror", "vo_create", "vo_destroy", "vo_self_check",
    "vo_grid_keypoints", "vo_anms", "vo_lk_track", "vo_debug_pyramid_level", "vo_debug_pyramid_padded", "vo_fmat_ransac", "vo_triangulate",
    "vo_pnp_ransac", "vo_debug_last_pnp", "vo_debug_last_fmat", "vo_debug_epnp", "vo_transform_points", "vo_bgr_to_gray", "vo_sor_cloud", "vo_pose_from_pnp",
    "vo_dense_lk_tracking", "vo_fmat_thresholding", "vo_stereo_triangulate", "vo_insert_keyframe",
    "vo_track_frame", "vo_pnp_frame", "vo_seq_init", "vo_seq_track", "vo_seq_prefetch", "vo_seq_announce", "vo_seq_get_reference", "vo_cuda_stream",
    "vo_sync", "vo_profile_enable", "vo_profile_read", "vo_debug_timeline", "vo_launch_count", "vo_lk_work", "vo_lk_slow_paths", "vo_measure_fp32_peak", "vo_measure_int32_peak",
    "vo_synth_render_dev", "vo_alloc_host", "vo_free_host", "vo_alloc_dev", "vo_free_dev", "vo_memcpy_d2h",
    "vo_memcpy_h2d",
    "vo_sgbm_default_params", "vo_sgbm_compute", "vo_stereo_match", "vo_reproject_disparity", "vo_sgbm_timing",
    "vo_debug_sgbm_stage", "vo_orb_describe", "vo_orb_smooth", "vo_orb_angles", "vo_orb_harris", "vo_fast9", "vo_orb_detect_and_compute",
]


def load():
    """Load libvo_b200.so; raises if it has not been built (python -m ros_stereo_slam_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libvo_b200.so is missing at %s -- build it with `python -m ros_stereo_slam_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for s in SYMBOLS:
        getattr(lib, s)
    lib.vo_last_error.restype = C.c_char_p
    lib.vo_strerror.restype = C.c_char_p
    lib.vo_strerror.argtypes = [C.c_int]
    lib.vo_cuda_stream.restype = C.c_void_p
    lib.vo_cuda_stream.argtypes = [C.c_void_p]
    lib.vo_launch_count.restype = C.c_int64
    lib.vo_launch_count.argtypes = [C.c_void_p]
    lib.vo_create.argtypes = [C.POINTER(VoParams), C.POINTER(C.c_void_p)]
    lib.vo_destroy.argtypes = [C.c_void_p]
    lib.vo_default_params.argtypes = [C.POINTER(VoParams)]
    lib.vo_default_params.restype = None
    lib.vo_sgbm_default_params.argtypes = [C.POINTER(VoSgbmParams)]
    lib.vo_sgbm_default_params.restype = None
    _lib = lib
    return lib


def _strerror(code):
    try:
        return load().vo_strerror(code).decode()
    except Exception:
        return "error"


def check(code, ok=(VO_OK,)):
    if code not in ok:
        raise VoError(code, load().vo_last_error().decode())
    return code


def default_params(**kw):
    p = VoParams()
    load().vo_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p
