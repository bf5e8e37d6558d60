"""Host-side logic of the multi-GPU bench path (replicas only), covered with
world_size-2 gloo on CPU: the aggregate is max(time) over ranks and sum(units) over ranks."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    ms, frames, kp = bench.aggregate(10.0 + 5.0 * rank, 100, 1000 * (rank + 1), world, backend="gloo")
    assert bench.rank_times(10.0 + 5.0 * rank, world, backend="gloo") == [10.0, 15.0]
    q.put((rank, ms, frames, kp))
    dist.barrier()
    dist.destroy_process_group()


def test_aggregate_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = [q.get(timeout=120) for _ in range(2)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, frames, kp in out:
        assert ms == 15.0 and frames == 200 and kp == 3000


def test_aggregate_single():
    import bench
    assert bench.aggregate(12.5, 10, 99, 1) == (12.5, 10, 99)
    assert bench.rank_times(12.5, 1) == [12.5]


def test_dist_env_defaults(monkeypatch):
    import bench
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    assert bench.dist_env() == (0, 1, 0)


def test_clock_sampler_without_nvml():
    """no GPU driver here: the sampler reports that instead of raising, and a rank that does not sample says so"""
    import bench
    s = bench.ClockSampler([0])
    s.start()
    s.mark_begin()
    s.mark_end()
    out = s.stop()
    assert out["sm_mhz"] is None and out["reasons"] and "NVML" in out["reasons"][0]
    idle = bench.ClockSampler([])
    assert idle.stop()["reasons"] == ["not sampled on this rank"]
