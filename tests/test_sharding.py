"""Multi-rank host logic on the CPU (gloo, world_size 2): frames shard by contiguous ranges, the only
exchange is the result gather, and the gathered records equal the unsharded run."""
import os
import sys

import numpy as np
import pytest

from trik_media_sensors_dsp_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_everything_once():
    for n in (0, 1, 7, 8, 1024, 4097):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = sharding.partition(n, world, r)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [sharding.partition(n, world, r)[1] - sharding.partition(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_streams_round_robin():
    assert sharding.streams_of_rank(10, 4, 1) == [1, 5, 9]
    assert sorted(sum((sharding.streams_of_rank(1024, 8, r) for r in range(8)), [])) == list(range(1024))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from oracle import ref
    from trik_media_sensors_dsp_b200 import sharding as sh, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, w, h = 11, 160, 120
    lo, hi = sh.partition(n, world, rank)
    orc = ref.OracleSensor("wl", w, h)                       # stand-in for the codec: the test is about the plumbing
    local = np.zeros((hi - lo, 16), dtype=np.uint8)
    for i in range(lo, hi):
        ok, out = orc.process(synth.make_frame("scene", i, w, h, "yuyv"), ref.RangeInArgs(0, 359, 0, 100, 0, 40, 0))
        local[i - lo] = np.frombuffer(ref.struct_bytes(out), dtype=np.uint8)
    full = sh.gather_records(local, n, world, rank)
    dist.barrier()
    if rank == 0:
        q.put(full.tobytes())
    dist.destroy_process_group()


def test_gloo_world2_gather_equals_unsharded():
    import torch.multiprocessing as mp
    from oracle import ref
    from trik_media_sensors_dsp_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, w, h = 11, 160, 120
    orc = ref.OracleSensor("wl", w, h)
    want = b""
    for i in range(n):
        ok, out = orc.process(synth.make_frame("scene", i, w, h, "yuyv"), ref.RangeInArgs(0, 359, 0, 100, 0, 40, 0))
        want += ref.struct_bytes(out)
    assert got == want
