#!/usr/bin/env python
"""bench.py -- frames/s of the VO front-end hot path on synthetic KITTI-shaped stereo sequences.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], made concrete in SURVEY.md section 8d): 1241x376 u8 stereo
sequence, grid step 5 -> 18,278 keypoints entering the stereo LK of every frame, 1024 PnP
hypotheses, keyframe inserted on every frame so that each step does KLT (temporal + stereo),
F-matrix RANSAC (x2), triangulation and PnP-RANSAC + refinement -- the metric's
"KLT + triangulate + PnP-RANSAC" per frame.  One step = one frame.

  value : frames/s with the frames already resident in HBM (device pointers), CUDA events on
          the library's stream, max over ranks.
  e2e   : the same through the C ABI with HOST buffers: every step copies that step's left and
          right image from pinned host memory (H2D inside the timed region) and reads the pose
          and counters back (D2H).  The caller announces frame n+1 with vo_seq_prefetch before
          it calls vo_seq_track for frame n, so that copy runs under frame n's processing.
  N > 1 : replicas only: one independent sequence per GPU, no data-path
          collective; torch.distributed (NCCL) is used for the barrier and the max/sum of the
          timings only.
  matrix: (N = 1) the other BASELINE.json configurations and the reference's real settings as named
          sub-objects of the same JSON line, each with its own oracle-parity boolean: config 1 (grid step 9,
          100 frames, reference defaults), the reference keyframe rule (inliers < 200) at config-2 sizes,
          3-channel BGR input, config 3 (115k candidates -> ANMS -> LK) and config 4 (PnP stress).
  --impl reference : the reference's own CPU implementation of the path (oracle.glue: the reference glue
          called through to cv2, all host threads) on the same config; frames come from the oracle's numpy
          renderer, no GPU and no repo library is touched.  With N > 1 rank 0 runs N independent sequences
          (one process each, the host cores divided between them) and reports their aggregate.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WIDTH, HEIGHT = 1241, 376
GRID_STEP = 5
GRID_KEYPOINTS = 18278
PNP_ITERS = 1024
WORKLOAD = "kitti00_synth_1241x376_grid5_18278kp_pnp1024_keyframe_every_frame"
METRIC = "frames/sec for KLT+triangulate+PnP-RANSAC at 1241x376; Mkeypoints/s tracked"
INT_MAX = 2 ** 31 - 1
TOL_RAD, TOL_M = 1e-4, 1e-3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=25, help="frames of the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-matrix", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    return ap.parse_args()


def bench_config(world, steps, warmup):
    """The workload description BOTH arms print (identical keys and values)."""
    return {"workload": WORKLOAD, "image": "1241x376 u8 c1", "grid_step": GRID_STEP, "grid_keypoints": GRID_KEYPOINTS,
            "pnp_iterations": PNP_ITERS, "keyframe_rule": "every frame", "frames": steps, "warmup_frames": warmup,
            "sequences": world, "scene_seed": "same scene on every rank (weak scaling: identical work per GPU)", "parallelism": "replicas x%d (one sequence per GPU, no collective)" % world}


# ----------------------------------------------------------------------------- helpers
class ClockSampler:
    """SM clocks and throttle reasons of the given GPUs, sampled every ~10 ms by an NVML thread of THIS process while it
    is running (started before the warm-up, so the timed region lies inside the sampled span; mark() brackets it).
    A spawned `nvidia-smi -lms` -- the first version -- starts up inside a 20 ms timed region and takes driver locks
    the frame loop needs; at N > 1 only rank 0 samples (all N GPUs), the other ranks run undisturbed."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, devices):
        self.devices = list(devices)
        self.rows = []          # (t, device, sm_mhz, max_mhz, reasons_bitmask)
        self.t0 = self.t1 = None
        self.stop_flag = False
        self.thread = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            ids = [int(x) for x in vis.split(",")] if vis and all(x.strip().isdigit() for x in vis.split(",")) else None
            hs = [pynvml.nvmlDeviceGetHandleByIndex(ids[d] if ids else d) for d in self.devices]
            mx = [pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM) for h in hs]
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        except Exception as e:       # noqa: BLE001
            self.err = "NVML unavailable: %s" % e
            return

        def loop():
            while not self.stop_flag:
                t = time.perf_counter()
                for d, h, m in zip(self.devices, hs, mx):
                    try:
                        self.rows.append((t, d, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), m,
                                          get_reasons(h)))
                    except Exception:   # noqa: BLE001
                        pass
                # dense at first (a 20-step timed region lasts ~20 ms), sparse once it is covered: every NVML query takes
                # driver locks the frame loop's copies need (a 200-step e2e leg lost 4 % to a fixed 10 ms period)
                dense = self.t0 is None or t - self.t0 < 0.06
                time.sleep(0.008 if dense else 0.1)

        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "not sampled on this rank"]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        inside = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or 1e30)]
        # a 20-step region lasts ~20 ms: if no sample fell inside it, the neighbouring ones (same load) stand in
        used = inside if inside else [r for r in self.rows if self.t0 is None or abs(r[0] - self.t0) < 0.25]
        sm = [r[2] for r in used]
        bits = 0
        for r in used:
            bits |= r[4]
        per_gpu = {}
        for r in used:
            per_gpu.setdefault(r[1], []).append(r[2])
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max((r[3] for r in used), default=None),
               "samples": len(used), "samples_inside_timed_region": len(inside),
               "reasons": [n for n, b in self.REASONS if bits & b], "source": "NVML thread, ~10 ms period over the first 60 ms of the timed region, 100 ms after"}
        if len(self.devices) > 1:
            out["sm_mhz_per_gpu"] = {str(d): float(np.median(v)) for d, v in sorted(per_gpu.items())}
        return out


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def aggregate(elapsed_ms, frames, keypoints, world, backend=None):
    """Max elapsed time over ranks, sum of units over ranks.  Host-side logic of the N>1 path
    (tested with gloo, world_size 2, in tests/test_bench_dist.py)."""
    if world == 1:
        return elapsed_ms, frames, keypoints
    import torch
    import torch.distributed as dist
    dev = "cuda" if (backend or dist.get_backend()) == "nccl" else "cpu"
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    u = torch.tensor([float(frames), float(keypoints)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), int(u[0].item()), int(u[1].item())


def rank_times(elapsed_ms, world, backend=None):
    """Every rank's elapsed time (ms), rank order: shows whether a sub-linear aggregate is one slow GPU (the aggregate
    divides by the MAX over ranks) or a slowdown of all of them."""
    if world == 1:
        return [float(elapsed_ms)]
    import torch
    import torch.distributed as dist
    dev = "cuda" if (backend or dist.get_backend()) == "nccl" else "cpu"
    mine = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    allt = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allt, mine)
    return [float(x.item()) for x in allt]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def pin_threads(rank, world):
    """Each rank's caller thread (and the library's worker thread it spawns) on its own slice of the host cores:
    with every rank free to run anywhere the frame loops of 8 ranks preempt each other."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(2, len(cores) // max(world, 1))
        mine = cores[(rank * per) % len(cores):(rank * per) % len(cores) + per]
        if len(mine) >= 2:
            os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None


# ----------------------------------------------------------------------------- device-resident sequences
class DeviceSequence:
    """nf stereo frames of scene `seed` rendered on the GPU (harness kernel) with a pinned host copy."""

    def __init__(self, fe, lib, seed, nf, channels=1):
        from ros_stereo_slam_b200 import _lib
        self.fe, self.lib, self._lib, self.nf, self.cn = fe, lib, _lib, nf, channels
        self.gray_bytes = WIDTH * HEIGHT
        self.img_bytes = self.gray_bytes * channels
        self.stride = WIDTH * channels
        self.d = C.c_void_p()
        self.h = C.c_void_p()
        _lib.check(lib.vo_alloc_dev(fe.h, C.byref(self.d), C.c_uint64(2 * nf * self.img_bytes)))
        _lib.check(lib.vo_alloc_host(C.byref(self.h), C.c_uint64(2 * nf * self.img_bytes)))
        if channels == 1:
            for i in range(nf):
                for eye in (0, 1):
                    _lib.check(lib.vo_synth_render_dev(fe.h, seed, i, eye, C.c_void_p(self.dptr(i, eye))))
            _lib.check(lib.vo_memcpy_d2h(fe.h, self.h, self.d, C.c_uint64(2 * nf * self.img_bytes)))
        else:
            # the reference's real input: a gray frame read by imread comes back as 3 equal channels
            arr = self.host_array()
            for i in range(nf):
                for eye in (0, 1):
                    g = fe.synth_render(seed, i, eye)
                    arr[i, eye] = g[:, :, None]
            _lib.check(lib.vo_memcpy_h2d(fe.h, self.d, self.h, C.c_uint64(2 * nf * self.img_bytes)))

    def dptr(self, i, eye):
        return self.d.value + (2 * i + eye) * self.img_bytes

    def hptr(self, i, eye):
        return self.h.value + (2 * i + eye) * self.img_bytes

    def host_array(self):
        buf = (C.c_uint8 * (2 * self.nf * self.img_bytes)).from_address(self.h.value)
        shape = (self.nf, 2, HEIGHT, WIDTH) if self.cn == 1 else (self.nf, 2, HEIGHT, WIDTH, self.cn)
        return np.frombuffer(buf, np.uint8).reshape(shape)

    def free(self):
        self.lib.vo_free_dev(self.fe.h, self.d)
        self.lib.vo_free_host(self.h)


def run_frames(fe, lib, seq, first, count, device_resident, record=None):
    """vo_seq_track over frames [first, first+count); returns keypoints that entered an LK launch."""
    from ros_stereo_slam_b200 import _lib
    res = _lib.VoFrameResult()
    kp = 0
    for i in range(first, first + count):
        if device_resident:
            if i + 1 < first + count:     # the next frame is resident too: announce it (look-ahead of its temporal LK)
                _lib.check(lib.vo_seq_announce(fe.h, C.c_void_p(seq.dptr(i + 1, 0)), C.c_void_p(seq.dptr(i + 1, 1)), seq.stride, 1))
            r = lib.vo_seq_track(fe.h, C.c_void_p(seq.dptr(i, 0)), C.c_void_p(seq.dptr(i, 1)), seq.stride, 1, 0, C.byref(res))
        else:
            if i + 1 < first + count:
                _lib.check(lib.vo_seq_prefetch(fe.h, C.c_void_p(seq.hptr(i + 1, 0)), C.c_void_p(seq.hptr(i + 1, 1)), seq.stride))
            r = lib.vo_seq_track(fe.h, C.c_void_p(seq.hptr(i, 0)), C.c_void_p(seq.hptr(i, 1)), seq.stride, 0, 0, C.byref(res))
        _lib.check(r)
        kp += res.n_lk_in + res.n_lk_in_stereo
        if record is not None:
            record.append(dict(n_lk_in=res.n_lk_in, n_tracked=res.n_tracked, n_inliers=res.n_inliers,
                               keyframe=bool(res.keyframe), n_kf_points=res.n_kf_points, rvec=np.array(res.rvec),
                               tvec=np.array(res.tvec)))
    return kp


def seq_init(fe, lib, seq, device_resident):
    from ros_stereo_slam_b200 import _lib
    n0 = C.c_int()
    _lib.check(lib.vo_seq_init(fe.h, C.c_void_p(seq.dptr(0, 0) if device_resident else seq.hptr(0, 0)),
                               C.c_void_p(seq.dptr(0, 1) if device_resident else seq.hptr(0, 1)), seq.stride,
                               1 if device_resident else 0, C.byref(n0)))
    return n0.value


def records_match(ours, ref):
    """Per-frame parity of a sequence against oracle.glue.run_sequence: counters identical, pose within the
    BASELINE.json tolerances (1e-4 rad, 1e-3 m)."""
    if len(ours) < len(ref) or not ref:
        return False
    for a, b in zip(ours, ref):
        if (a["n_lk_in"], a["n_tracked"], a["n_inliers"], a["keyframe"]) != (b["n_lk_in"], b["n_tracked"], b["n_inliers"], b["keyframe"]):
            return False
        if b["keyframe"] and a["n_kf_points"] != b.get("n_kf_points", a["n_kf_points"]):
            return False
        if np.abs(a["rvec"] - b["rvec"]).max() > TOL_RAD or np.abs(a["tvec"] - b["tvec"]).max() > TOL_M:
            return False
    return True


def timed_sequence(fe, lib, seq, warm, count, device_resident=True):
    """(ms, keypoints, records) of `count` frames after `warm` untimed ones, CUDA events on the library stream."""
    import torch
    stream = torch.cuda.ExternalStream(lib.vo_cuda_stream(fe.h))
    seq_init(fe, lib, seq, device_resident)
    rec = []
    run_frames(fe, lib, seq, 1, warm, device_resident, rec)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    kp = run_frames(fe, lib, seq, 1 + warm, count, device_resident, rec)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), kp, rec


# ----------------------------------------------------------------------------- the BASELINE matrix (N = 1)
def matrix(args, dev):
    """BASELINE.json configs 1, 3, 4, the reference keyframe rule and 3-channel input, each beside its oracle."""
    import cv2
    from oracle import glue, synth
    from ros_stereo_slam_b200 import VisualFrontEnd
    out = {}
    cores = len(os.sched_getaffinity(0))
    cv2.setNumThreads(cores)

    def seq_case(name, nf, warm, parity_frames, channels=1, **params):
        fe = VisualFrontEnd(device=dev, channels=channels, **params)
        seq = DeviceSequence(fe, fe.lib, args.seed, nf, channels)
        ms, kp, rec = timed_sequence(fe, fe.lib, seq, warm, nf - 1 - warm, True)
        ms_h, _, _ = timed_sequence(fe, fe.lib, seq, warm, nf - 1 - warm, False)
        arr = seq.host_array()
        n = parity_frames + 1
        t0 = time.perf_counter()
        ref = glue.run_sequence([arr[i, 0] for i in range(n)], [arr[i, 1] for i in range(n)], step=params["grid_step"],
                                pnp_iters=params["pnp_iters"], kf_min_inliers=min(params["kf_min_inliers"], 10 ** 9))
        t_cpu = time.perf_counter() - t0
        frames = nf - 1 - warm
        o = {"frames": frames, "frames_per_s": round(frames / (ms * 1e-3), 2), "ms_per_frame": round(ms / frames, 4),
             "e2e_frames_per_s": round(frames / (ms_h * 1e-3), 2),
             "keypoints_in_first_frame": rec[0]["n_lk_in"], "keyframes": int(sum(r["keyframe"] for r in rec[warm:])),
             "last_frame": {k: rec[-1][k] for k in ("n_lk_in", "n_tracked", "n_inliers")},
             "oracle_frames": len(ref), "oracle_parity": bool(records_match(rec, ref)),
             "cpu_frames_per_s": round(len([r for r in ref if "ms" in r]) / max(sum(r.get("ms", 0) for r in ref) * 1e-3, 1e-9), 3),
             "cpu_wall_s": round(t_cpu, 2), "params": {k: (v if v != INT_MAX else "INT_MAX") for k, v in params.items()},
             "channels": channels}
        seq.free()
        fe.close()
        out[name] = o

    # config 1: 100 frames, grid step 9 (5,440 keypoints), every reference default (PnP 100 iterations,
    # keyframe when inliers < 200, OpenCV's adaptive RANSAC stop)
    seq_case("config1_step9_100frames_reference_defaults", 101 + 3, 3, 10, grid_step=9, pnp_iters=100, kf_min_inliers=200,
             ransac_exhaustive=0)
    # the reference's keyframe rule (src/VisualSLAM.cpp:120) at the headline sizes: the keyframe is discovered
    # after PnP, so its stereo chain runs serially behind the tracking chain
    seq_case("config2_sizes_keyframe_rule_inliers_lt_200", 100 + 3, 3, 6, grid_step=GRID_STEP, pnp_iters=PNP_ITERS,
             kf_min_inliers=200, ransac_exhaustive=0)
    # the reference's real input: 3-channel frames (imread default), headline settings
    seq_case("bgr_3channel_headline_settings", 24 + 3, 3, 2, channels=3, grid_step=GRID_STEP, pnp_iters=PNP_ITERS,
             kf_min_inliers=INT_MAX, ransac_exhaustive=1)

    # config 3: step-2 candidates (115,134) -> ANMS(80,000) -> 4-level 21x21 LK
    fe = VisualFrontEnd(device=dev, max_points=131072)
    sc = synth.Scene(2)
    L0, L1 = sc.render(0, "L"), sc.render(1, "L")
    cand = glue.dense_keypoint_extractor(HEIGHT, WIDTH, 2)
    gx = cv2.Sobel(L0, cv2.CV_32F, 1, 0, ksize=3)
    gy = cv2.Sobel(L0, cv2.CV_32F, 0, 1, ksize=3)
    resp = cv2.boxFilter(gx * gx + gy * gy, -1, (7, 7))[cand[:, 1].astype(int), cand[:, 0].astype(int)].astype(np.float32)
    keep = fe.adaptiveNonMaximalSuppresion(cand, resp, 80000)
    t0 = time.perf_counter()
    keep = fe.adaptiveNonMaximalSuppresion(cand, resp, 80000)
    t_anms = time.perf_counter() - t0
    pts = cand[keep]
    fe.calcOpticalFlowPyrLK(L0, L1, pts)
    fe.profile_enable(["lk"]); fe.profile_read(reset=True)
    for _ in range(5):
        p, st, _e = fe.calcOpticalFlowPyrLK(L0, L1, pts)
    l, ms = fe.profile_read(reset=True)["lk"]
    fe.profile_enable(None)
    t0 = time.perf_counter()
    p0, st0, _ = cv2.calcOpticalFlowPyrLK(L0, L1, pts.reshape(-1, 1, 2), None)
    t_cv = time.perf_counter() - t0
    st0 = st0.ravel(); p0 = p0.reshape(-1, 2)
    sub = np.arange(0, len(cand), 23)
    anms_ok = np.array_equal(fe.adaptiveNonMaximalSuppresion(cand[sub], resp[sub], 3000), glue.anms(cand[sub], resp[sub], 3000))
    out["config3_density_115k_anms_80k_lk"] = {
        "candidates": int(len(cand)), "kept": int(len(keep)), "anms_call_ms": round(t_anms * 1e3, 2),
        "lk_kernel_ms": round(ms / l, 4), "mkeypoints_per_s": round(len(pts) / (ms / l * 1e-3) / 1e6, 2),
        "cv2_lk_ms": round(t_cv * 1e3, 1), "cores": cores,
        "oracle_parity": bool(np.array_equal(st, st0) and np.array_equal(p[st0 == 1], p0[st0 == 1]) and anms_ok),
        "parity_definition": "status identical and every tracked position bit-identical to cv2.calcOpticalFlowPyrLK; "
                             "ANMS kept set identical to oracle.glue.anms on a 1/23 subsample"}
    fe.close()

    # config 4: PnP-RANSAC stress, N = 20,000, 50 % outliers, 4096 iterations, LM refinement on the inliers
    X, xy, _, _, _ = synth.pnp_stress_case(20000, 0.5, 0.3, seed=3)
    t0 = time.perf_counter()
    ok, rv0, tv0, inl0 = cv2.solvePnPRansac(X.reshape(-1, 1, 3), xy.reshape(-1, 1, 2), glue.K, np.zeros((4, 1)), None, None,
                                            False, 4096, 1.0, 0.99)
    t_cv = time.perf_counter() - t0
    c4 = {"points": 20000, "outlier_fraction": 0.5, "iterations": 4096, "cv2_ms": round(t_cv * 1e3, 2), "cores": cores}
    for ex in (0, 1):
        f = VisualFrontEnd(device=dev, ransac_exhaustive=ex)
        f.solvePnPRansac(X, xy, 4096, 1.0, 0.99)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            r = f.solvePnPRansac(X, xy, 4096, 1.0, 0.99)
            ts.append(time.perf_counter() - t0)
        same = (np.array_equal(r["inliers"], inl0.ravel()) and np.abs(r["rvec"] - rv0.ravel()).max() <= TOL_RAD
                and np.abs(r["tvec"] - tv0.ravel()).max() <= TOL_M)
        c4["early_exit" if ex == 0 else "exhaustive"] = {
            "call_ms": round(min(ts) * 1e3, 3), "hypotheses_evaluated": int(len(f.last_pnp()["counts"])),
            "inliers": int(len(r["inliers"])), "oracle_parity": bool(same)}
        f.close()
    c4["parity_definition"] = ("inlier index set identical to cv2.solvePnPRansac, pose within 1e-4 rad / 1e-3 m (the exhaustive "
                               "run evaluates all 4096 hypotheses and must select the same record-setting model)")
    out["config4_pnp_stress_20k_50pct_4096"] = c4
    return out


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    from ros_stereo_slam_b200 import VisualFrontEnd, _lib
    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = local if world > 1 else 0
    torch.cuda.set_device(dev)
    pinned = pin_threads(local, world) if world > 1 else None
    K, W = args.steps, max(args.warmup, 3)
    nf = K + W + 1
    fe = VisualFrontEnd(device=dev, grid_step=GRID_STEP, pnp_iters=PNP_ITERS, kf_min_inliers=INT_MAX,
                        ransac_exhaustive=1)
    lib = fe.lib
    img_bytes = WIDTH * HEIGHT
    seed = args.seed                   # one independent replica of the SAME sequence per GPU: per-GPU work is fixed as N
                                       # grows (different scenes differ by +-4 % in work and the aggregate takes the slowest)
    seq = DeviceSequence(fe, lib, seed, nf)
    stream = torch.cuda.ExternalStream(lib.vo_cuda_stream(fe.h), device=torch.device("cuda", dev))

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(device_resident):
        import gc
        sampler = ClockSampler(range(world) if rank == 0 else [])
        if rank == 0:
            sampler.start()
        barrier()     # the first collective sets up the NCCL communicator: keep that out of the timed region's barrier
        seq_init(fe, lib, seq, device_resident)
        run_frames(fe, lib, seq, 1, W, device_resident)           # warm-up
        fe.profile_enable(["lk"])                                  # events around the LK launches only
        fe.profile_read(reset=True)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        gc.collect()
        gc.disable()                                               # no collection pause inside a 20 ms region
        barrier()
        sampler.mark_begin()
        l0 = fe.launch_count()
        t0 = time.perf_counter()
        w0 = fe.lk_work()
        e0.record(stream)
        rec = []
        kp = run_frames(fe, lib, seq, 1 + W, K, device_resident, rec)
        e1.record(stream)
        barrier()
        sampler.mark_end()
        gc.enable()
        wall_ms = (time.perf_counter() - t0) * 1e3
        launches = fe.launch_count() - l0
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        prof = fe.profile_read(reset=True)
        fe.profile_enable(None)
        w1 = fe.lk_work()
        return dict(ms=ms, wall_ms=wall_ms, kp=kp, launches=launches, clocks=clocks, lk=prof["lk"], rec=rec,
                    lk_work=(w1[0] - w0[0], w1[1] - w0[1]))

    dev_run = timed(True)
    host_run = timed(False)

    # ---- per-stage breakdown (separate, untimed pass with every kernel family bracketed)
    fe.profile_enable("all")
    fe.profile_read(reset=True)
    seq_init(fe, lib, seq, True)
    nb = min(20, K)
    run_frames(fe, lib, seq, 1, nb, True)
    breakdown = {k: {"launches": v[0] / nb, "ms_per_frame": v[1] / nb} for k, v in fe.profile_read(reset=True).items()}
    fe.profile_enable(None)

    int_peak = fe.measure_int32_peak()
    fp32_peak = fe.measure_fp32_peak()

    e_ms, frames, kps = aggregate(dev_run["ms"], K, dev_run["kp"], world)
    h_ms, frames_h, kps_h = aggregate(host_run["ms"], K, host_run["kp"], world)
    per_rank = [rank_times(dev_run["ms"], world), rank_times(host_run["ms"], world)]

    out = None
    if rank == 0:
        hbm_peak, hbm_src = load_peaks()
        fps = frames / (e_ms * 1e-3)
        fps_e2e = frames_h / (h_ms * 1e-3)
        lk_launches, lk_ms = dev_run["lk"]
        lk_avg_ms = lk_ms / max(lk_launches, 1)
        # algorithmic work: 441*(30 + 13*iters) integer ops per (point, level), with the
        # (point, level) pairs and iterations counted by the kernel over the timed region;
        # bytes per launch = both u8 pyramids + the int16x2 derivative pyramid + 21 B/point
        pl, it = dev_run["lk_work"]
        ops_per_launch = 441.0 * (30.0 * pl + 13.0 * it) / max(lk_launches, 1)
        kp_per_launch = dev_run["kp"] / max(lk_launches, 1)
        lk_bytes = 2 * 619930 + 4 * 619930 + 21.0 * kp_per_launch
        achieved_tops = ops_per_launch / (lk_avg_ms * 1e-3) / 1e12
        roofline = {
            "kernel": "lk_kernel (pyramidal LK, warp per keypoint, window sums in OpenCV's float order)",
            "bound": "int32_issue",
            "achieved": round(achieved_tops, 3), "peak": round(int_peak, 2), "unit": "Tops/s (integer)",
            "frac": round(achieved_tops / int_peak, 4),
            "peak_source": "IMAD microbenchmark measured in this run (vo_measure_int32_peak; multiply-add = 2 ops); the "
                           "kernel issues DP2A/IMAD/SHF, MEASURED_PEAKS.json has no integer figure; FFMA peak measured in "
                           "this run for context: %.2f TFLOP/s" % fp32_peak,
            "avg_launch_ms": round(lk_avg_ms, 4), "launches_timed": lk_launches,
            "avg_launch_ms_note": "two LK launches (temporal, stereo) run concurrently on two streams inside a step; the "
                                  "event brackets include that overlap -- stand-alone launch time: profiles/",
            "ops_per_launch": ops_per_launch, "keypoints_per_launch": round(kp_per_launch, 1),
            "iterations_per_point_level": round(it / max(pl, 1), 2),
            "ops_model": "441*(30*point_levels + 13*iterations), counted by the kernel",
            "hbm": {"achieved": round(lk_bytes / (lk_avg_ms * 1e-3) / 1e9, 2), "peak": hbm_peak, "unit": "GB/s",
                    "frac": round(lk_bytes / (lk_avg_ms * 1e-3) / 1e9 / hbm_peak, 5), "peak_source": hbm_src,
                    "note": "working set (~5 MB of pyramids) is L2-resident; HBM is not the bound"},
            "traffic": None,
            "traffic_note": "dram bytes per launch from the ncu --set full capture of this kernel: profiles/ (r02_*)",
            "algorithmic_bytes_per_launch": round(lk_bytes),
        }
        stage_sum = sum(v["ms_per_frame"] for v in breakdown.values())
        out = {
            "metric": METRIC,
            "value": round(fps, 2), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(e_ms / K, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/int32 fixed-point (LK) + f64 (RANSAC solvers)", "data": "synthetic",
            "mkeypoints_per_s": round(kps / (e_ms * 1e-3) / 1e6, 3),
            "keypoints_per_frame": round(kps / frames, 1),
            "config": bench_config(world, K, W),
            "notes": {"ransac_exhaustive": "all %d PnP hypotheses are evaluated (the reference arm keeps OpenCV's adaptive "
                                           "early exit, i.e. does less work)" % PNP_ITERS,
                      "keyframe_rule": "kf_min_inliers=INT_MAX: a keyframe (stereo LK + F-RANSAC + triangulation) on every frame",
                      "l2": "every step reads two frames not touched before; the %d-frame set is %.0f MB"
                            % (nf, 2 * nf * img_bytes / 1e6),
                      "timing": "CUDA events on the library stream, barrier+synchronize both sides, max over ranks",
                      "renderer": "GPU harness kernel (vo_synth_render_dev); the reference arm renders the same scene with "
                                  "oracle/synth.py (tests/test_gpu_stages.py pins the two against each other)",
                      "host_threads": pinned},
            "e2e": {"value": round(fps_e2e, 2), "unit": "frames/s", "ms_per_step": round(h_ms / K, 4),
                    "h2d_bytes_per_step": 2 * img_bytes, "d2h_bytes_per_step": C.sizeof(_lib.VoFrameResult),
                    "api": "vo_seq_prefetch(next host left/right) + vo_seq_track(host left, host right) -> "
                           "vo_frame_result; every frame's H2D copy (pinned memory) and result read-back happen "
                           "inside the timed region, the copy of frame n+1 overlapping the processing of frame n"},
            "gpu_launches": int(dev_run["launches"]),
            "gpu_launches_per_step": round(dev_run["launches"] / K, 1),
            "clocks": dev_run["clocks"],
            "wall_ms_per_step": round(dev_run["wall_ms"] / K, 4),
            "rank_ms_per_step": {"value": [round(x / K, 4) for x in per_rank[0]], "e2e": [round(x / K, 4) for x in per_rank[1]],
                                 "note": "every rank's own device-timed ms per step; the aggregate uses the slowest"},
            "roofline": roofline,
            "stage_ms_per_frame_overlapped": {k: round(v["ms_per_frame"], 4) for k, v in breakdown.items()},
            "stage_ms_note": "per-launch event brackets from a separate profiling pass; the tracking and the stereo chain run "
                             "concurrently on two streams, so the brackets include queueing behind the other chain and their "
                             "sum (%.3f ms) exceeds the frame time (%.3f ms)" % (stage_sum, e_ms / K),
            "last_frame": {k: dev_run["rec"][-1][k] for k in ("n_lk_in", "n_tracked", "n_inliers", "n_kf_points")},
        }
        if world == 1 and not args.no_cpu_baseline:
            ncpu = min(args.cpu_frames, nf - 1) + 1
            arr = seq.host_array()
            Ls, Rs = [arr[i, 0] for i in range(ncpu)], [arr[i, 1] for i in range(ncpu)]
            out["cpu_baseline"] = cpu_reference(Ls, Rs, ncpu - 1)
            # parity of the timed sequence itself: the GPU records of the first frames against the oracle's
            out["oracle_parity"] = bool(records_match(sequence_records(fe, lib, seq, 4), cpu_records(Ls[:5], Rs[:5])))
            out["oracle_parity_definition"] = ("first 4 frames of the timed sequence: keypoints in, tracked, PnP inliers, keyframe "
                                               "points identical to oracle.glue.run_sequence (cv2), pose within 1e-4 rad / 1e-3 m")
            # side figure, not part of the metric: the dense-stereo path on the
            # first stereo pair of the sequence, through vo_sgbm_compute with host buffers, beside cv2 on the host
            out["dense_stereo"] = dense_stereo_side_figure(fe, arr[0, 0].copy(), arr[0, 1].copy())
            # and the loop detector's per-frame feature extraction
            out["loop_detector_orb"] = orb_side_figure(fe, arr[0, 0].copy())
    seq.free()
    fe.close()
    if out is not None and world == 1 and not args.no_matrix:
        out["matrix"] = matrix(args, dev)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


def sequence_records(fe, lib, seq, n):
    rec = []
    seq_init(fe, lib, seq, True)
    run_frames(fe, lib, seq, 1, n, True, rec)
    return rec


def cpu_records(Ls, Rs):
    from oracle import glue
    return glue.run_sequence(Ls, Rs, step=GRID_STEP, pnp_iters=PNP_ITERS, kf_min_inliers=10 ** 9)


def dense_stereo_side_figure(fe, L, R, reps=100):
    import cv2
    ref = cv2.StereoSGBM_create(1, 96, 7, 24, 96, 0, 60, 0, 3000, 5)      # reference src/StereoCV.cpp:39-50
    t0 = time.perf_counter()
    want = ref.compute(L, R)
    t_cpu = time.perf_counter() - t0
    got = fe.stereoMatch(L, R)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fe.stereoMatch(L, R)
        ts.append(time.perf_counter() - t0)
    return {"api": "vo_sgbm_compute(host left, host right) -> host int16 disparity, StereoSGBM::create(1, 96, 7, 24, 96, 0, "
                   "60, 0, 3000, 5)", "ms_per_pair": round(1e3 * float(np.median(ts[reps // 2:])), 4),
            "cv2_ms_per_pair": round(1e3 * t_cpu, 1), "cores": len(os.sched_getaffinity(0)),
            "bit_identical_to_cv2": bool(np.array_equal(got, want)),
            "device_pipeline_ms": round(fe.sgbm_timing()["pipeline"], 4),
            "note": "a repeated (size, parameters) replays one CUDA graph; per-stage times: tools/sgbm_bench.py"}


def orb_side_figure(fe, img, reps=60):
    import cv2
    ref = cv2.ORB_create()                                   # reference src/optimizationStuff.cpp:49
    t0 = time.perf_counter()
    kps, want = ref.detectAndCompute(img, None)
    t_cpu = time.perf_counter() - t0
    got = fe.orbDetectAndCompute(img, 500)
    key = sorted(range(len(kps)), key=lambda i: (kps[i].octave, kps[i].pt[1], kps[i].pt[0]))
    same = len(kps) == len(got["xy"]) and all(
        kps[i].pt == (float(got["xy"][j, 0]), float(got["xy"][j, 1])) and np.array_equal(want[i], got["desc"][j])
        for j, i in enumerate(key))
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fe.orbDetectAndCompute(img, 500)
        ts.append(time.perf_counter() - t0)
    return {"api": "vo_orb_detect_and_compute(host image, nfeatures 500) -> keypoints + 32-byte descriptors, = ORB::create()"
                   "->detectAndCompute", "ms_per_frame": round(1e3 * float(np.median(ts[reps // 2:])), 4),
            "cv2_ms_per_frame": round(1e3 * t_cpu, 2), "cores": len(os.sched_getaffinity(0)), "keypoints": len(kps),
            "bit_identical_to_cv2": bool(same)}


# ----------------------------------------------------------------------------- CPU reference
def cpu_stage_breakdown(L0, R0, L1, threads):
    """Per-stage ms of one frame of the reference path at the headline sizes: best of 3."""
    import cv2
    from oracle import glue
    cv2.setNumThreads(threads)
    pts = glue.dense_keypoint_extractor(HEIGHT, WIDTH, GRID_STEP)
    dist = np.zeros((4, 1))
    P1, P2 = glue.projection_matrices()

    def best(f, n=3):
        ts, r = [], None
        for _ in range(n):
            t0 = time.perf_counter()
            r = f()
            ts.append(time.perf_counter() - t0)
        return min(ts) * 1e3, r

    st = {}
    st["stereo_pyramid_lk"], (trk, status, _e) = best(lambda: cv2.calcOpticalFlowPyrLK(L0, R0, pts.reshape(-1, 1, 2), None))
    keep = status.ravel() == 1
    a, b = pts[keep], trk.reshape(-1, 2)[keep]
    st["stereo_fmat_ransac"], (F, mask) = best(lambda: cv2.findFundamentalMat(a, b, cv2.FM_RANSAC, 3.0, 0.99))
    m = mask.ravel() == 1
    a, b = a[m], b[m]
    st["triangulation"], xyz = best(lambda: glue.triangulate(P1, P2, a, b))
    st["temporal_pyramid_lk"], (trk, status, _e) = best(lambda: cv2.calcOpticalFlowPyrLK(L0, L1, a.reshape(-1, 1, 2), None))
    keep = status.ravel() == 1
    r2, r3, t2 = a[keep], xyz[keep], trk.reshape(-1, 2)[keep]
    st["temporal_fmat_ransac"], (F, mask) = best(lambda: cv2.findFundamentalMat(r2, t2, 8, 1.0, 0.99))
    m = mask.ravel() == 1
    t2, r3 = t2[m], r3[m]
    st["pnp_ransac_refine"], _r = best(lambda: cv2.solvePnPRansac(r3.reshape(-1, 1, 3), t2.reshape(-1, 1, 2), glue.K, dist, None,
                                                                 None, False, PNP_ITERS, 1.0, 0.99))
    return {k: round(v, 2) for k, v in st.items()}


def cpu_reference(Ls, Rs, n_frames, warm=1, threads=None, stages=True):
    """The reference's CPU path (oracle.glue: the reference glue over OpenCV) on a bounded sample
    of the same workload: grid step 5, PnP 1024 iterations, keyframe every frame."""
    import cv2
    from oracle import glue
    cores = len(os.sched_getaffinity(0))
    threads = threads or cores
    cv2.setNumThreads(threads)
    glue.run_sequence(Ls[:warm + 1], Rs[:warm + 1], step=GRID_STEP, pnp_iters=PNP_ITERS, kf_min_inliers=10 ** 9)
    t0 = time.perf_counter()
    recs = glue.run_sequence(Ls[:n_frames + 1], Rs[:n_frames + 1], step=GRID_STEP, pnp_iters=PNP_ITERS,
                             kf_min_inliers=10 ** 9)
    dt = time.perf_counter() - t0
    # run_sequence also performs the initial stereoTriangulate of frame 0; count per-frame time only
    per_frame = [r["ms"] for r in recs if "ms" in r]
    fps = 1e3 / (sum(per_frame) / max(len(per_frame), 1)) if per_frame else 0.0
    kp = sum(r["n_lk_in"] for r in recs) + len(per_frame) * GRID_KEYPOINTS
    out = {"value": round(fps, 3), "unit": "frames/s", "cores": threads, "kind": "port",
           "sample": "first %d frames of the same sequence (oracle.glue = reference glue restated over cv2 %s, "
                     "cv2.setNumThreads(%d) of %d host cores); wall %.1f s" % (len(per_frame), cv2.__version__, threads, cores, dt),
           "ms_per_frame": round(sum(per_frame) / max(len(per_frame), 1), 2),
           "mkeypoints_per_s": round(kp / (sum(per_frame) * 1e-3) / 1e6, 4) if per_frame else 0.0}
    if stages:
        # the 1-thread figure and the per-stage split (only LK/pyrDown are multithreaded in OpenCV)
        n1 = min(3, n_frames)
        cv2.setNumThreads(1)
        r1 = glue.run_sequence(Ls[:n1 + 1], Rs[:n1 + 1], step=GRID_STEP, pnp_iters=PNP_ITERS, kf_min_inliers=10 ** 9)
        pf1 = [r["ms"] for r in r1 if "ms" in r]
        out["threads_1"] = {"value": round(1e3 / (sum(pf1) / max(len(pf1), 1)), 3), "unit": "frames/s", "threads": 1,
                            "ms_per_frame": round(sum(pf1) / max(len(pf1), 1), 2), "frames": len(pf1)}
        out["stage_ms"] = {"threads_%d" % threads: cpu_stage_breakdown(Ls[0], Rs[0], Ls[1], threads),
                           "threads_1": cpu_stage_breakdown(Ls[0], Rs[0], Ls[1], 1)}
        out["cpu_features"] = [l.strip() for l in cv2.getBuildInformation().splitlines() if "Baseline:" in l or "Dispatched code" in l]
        cv2.setNumThreads(threads)
    return out


def _render_pair(job):
    seed, i = job
    from oracle import synth
    sc = synth.Scene(seed)
    return sc.render(i, "L"), sc.render(i, "R")


def render_sequence_cpu(seed, nf, workers):
    """nf stereo frames of scene `seed` with the oracle's numpy renderer (no GPU, no repo library)."""
    from concurrent.futures import ProcessPoolExecutor
    jobs = [(seed, i) for i in range(nf)]
    if workers > 1:
        with ProcessPoolExecutor(max_workers=workers) as ex:
            pairs = list(ex.map(_render_pair, jobs))
    else:
        pairs = [_render_pair(j) for j in jobs]
    return [p[0] for p in pairs], [p[1] for p in pairs]


def _reference_replica(job):
    """One independent sequence on a slice of the host cores (N > 1: one of N concurrent replicas)."""
    seed, nf, warm, steps, cores, q_ready, q_go = job
    import cv2
    try:
        os.sched_setaffinity(0, cores)
    except Exception:
        pass
    cv2.setNumThreads(len(cores))
    from oracle import glue
    Ls, Rs = render_sequence_cpu(seed, nf, 1)
    # warm-up frames (untimed), then exactly `steps` timed frames
    state = glue.run_sequence(Ls[:warm + 1], Rs[:warm + 1], step=GRID_STEP, pnp_iters=PNP_ITERS, kf_min_inliers=10 ** 9)
    q_ready.put(1)
    q_go.get()
    t0 = time.perf_counter()
    recs = glue.run_sequence(Ls[warm:], Rs[warm:], step=GRID_STEP, pnp_iters=PNP_ITERS, kf_min_inliers=10 ** 9)
    dt = time.perf_counter() - t0
    per_frame = [r["ms"] for r in recs if "ms" in r]
    return dict(frames=len(per_frame), ms=sum(per_frame), wall=dt, kp=sum(r["n_lk_in"] for r in recs) + len(per_frame) * GRID_KEYPOINTS,
                n_state=len(state))


def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return
    import cv2
    from oracle import glue
    K, W = args.steps, max(args.warmup, 3)
    n_seq = max(args.gpus, 1)
    cores = sorted(os.sched_getaffinity(0))
    nf = K + W + 1
    if n_seq == 1:
        Ls, Rs = render_sequence_cpu(args.seed, nf, min(len(cores), 32))
        cv2.setNumThreads(len(cores))
        glue.run_sequence(Ls[:W + 1], Rs[:W + 1], step=GRID_STEP, pnp_iters=PNP_ITERS, kf_min_inliers=10 ** 9)   # W warm-up frames
        t0 = time.perf_counter()
        # the timed frames continue from frame W: its stereo initialisation is outside the per-frame times
        recs = glue.run_sequence(Ls[W:], Rs[W:], step=GRID_STEP, pnp_iters=PNP_ITERS, kf_min_inliers=10 ** 9)
        wall = time.perf_counter() - t0
        per_frame = [r["ms"] for r in recs if "ms" in r]
        frames, ms = len(per_frame), sum(per_frame)
        kp = sum(r["n_lk_in"] for r in recs) + frames * GRID_KEYPOINTS
        fps = frames / (ms * 1e-3)
        threads = len(cores)
        sample = ("%d timed frames after %d warm-up frames of the config's sequence, one process, cv2.setNumThreads(%d); "
                  "wall %.1f s" % (frames, W, threads, wall))
    else:
        # N independent sequences, one process each, the host cores divided between them, started together
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        per = max(1, len(cores) // n_seq)
        mgr = ctx.Manager()
        q_ready, q_go = mgr.Queue(), mgr.Queue()
        jobs = [(args.seed, nf, W, K, cores[i * per:(i + 1) * per] or cores[-per:], q_ready, q_go) for i in range(n_seq)]
        with ctx.Pool(n_seq) as pool:
            res = pool.map_async(_reference_replica, jobs)
            for _ in range(n_seq):
                q_ready.get()
            for _ in range(n_seq):
                q_go.put(1)
            outs = res.get()
        frames = sum(o["frames"] for o in outs)
        ms = max(o["ms"] for o in outs)          # max over replicas, like max over ranks
        kp = sum(o["kp"] for o in outs)
        fps = frames / (ms * 1e-3)
        threads = per * n_seq
        sample = ("%d concurrent independent sequences x %d timed frames (one process each, %d cv2 threads each on disjoint "
                  "cores); aggregate frames / slowest replica's time" % (n_seq, K, per))
    cb = {"value": round(fps, 3), "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample,
          "ms_per_frame": round(ms / max(frames, 1) * n_seq, 2), "cv2": cv2.__version__}
    out = {
        "impl": "reference",
        "metric": METRIC,
        "value": cb["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": round(ms / max(frames, 1) * n_seq, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int16 fixed-point (LK) + f64 (RANSAC solvers)", "data": "synthetic",
        "mkeypoints_per_s": round(kp / (ms * 1e-3) / 1e6, 4),
        "config": bench_config(n_seq, K, W),
        "notes": {"renderer": "oracle/synth.py (numpy); same scene, seeds and trajectory as the GPU harness renderer",
                  "ransac": "OpenCV's adaptive early exit (the reference's semantics)"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


class StdoutGuard:
    """Everything libraries print to fd 1 during the run (e.g. NCCL's version banner) is sent to
    stderr, so that the ONE JSON line is the only thing on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


if __name__ == "__main__":
    a = parse()
    result = []
    _print = print

    def capture(line):
        result.append(line)

    import builtins
    with StdoutGuard():
        builtins.print = capture
        try:
            if a.impl == "reference":
                run_reference(a)
            else:
                run_ours(a)
        finally:
            builtins.print = _print
    for line in result:
        _print(line, flush=True)
