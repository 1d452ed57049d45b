"""trikb200_processBatchMulti: one batch cut into contiguous frame ranges over several handles -- one per
GPU on a multi-GPU box, several on one device here when only one is visible (the splitting, the host threads and the
carried-state walk are the same code).  Results must be those of ONE handle taking the whole batch, i.e. of n sequential
process() calls, including the ov7670 line sensor's lagging band and the object sensor's persisting range."""
import os
import subprocess

import numpy as np
import pytest

from trik_media_sensors_dsp_b200 import build, open_sensor, sensors, synth, xdm

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(kind, n):
    if kind == "oo":
        arr = (xdm.ObjInArgsAlg * n)()
        for i in range(n):                       # the range is set at frames 0 and 7 only; everything else lives on the carried one
            arr[i] = xdm.ObjInArgsAlg(1, 0, 20, 80, 20, 50, 30, 0) if i == 0 else \
                xdm.ObjInArgsAlg(1, 0, 30, 70, 25, 60, 30, 0) if i == 7 else xdm.ObjInArgsAlg(0, 0, 0, 0, 0, 0, 0, 0)
        return "blobs", arr, 24
    if kind == "om":
        return "grid", xdm.MxnInArgsAlg(3, 3), 36
    if kind == "wo":
        return "scene", xdm.RangeInArgsAlg(300, 40, 20, 100, 30, 100, 0), 3
    return "scene", xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 40, 0), 3


@pytest.mark.parametrize("kind", ["wl", "wo", "ol", "oo", "om"])
@pytest.mark.parametrize("handles,n", [(2, 37), (3, 2), (4, 64)])
def test_ranges_over_handles_equal_one_handle(kind, handles, n):
    w, h = 320, 240
    ndev = sensors.lib().trikb200_deviceCount()
    fam, ia, nb = _case(kind, n)
    frames = synth.make_batch(fam, range(n), w, h, sensors.layout_of(xdm.KIND_OF[kind]))
    one = open_sensor(kind, w, h)
    ret, want = one.process_batch(frames, ia)
    assert ret == 0, sensors.last_error()
    codecs = [open_sensor(kind, w, h, device=d % ndev) for d in range(handles)]
    ret, got = codecs[0].process_batch(frames, ia, multi=codecs)
    assert ret == 0, sensors.last_error()
    for i in range(n):
        assert bytes(memoryview(got[i]))[:nb] == bytes(memoryview(want[i]))[:nb], (kind, i)
    # the state every handle is left with is the one after the last frame: one more frame through each
    if kind in ("ol", "oo"):
        nxt = synth.make_batch(fam, [n + 1], w, h, sensors.layout_of(xdm.KIND_OF[kind]))
        follow = xdm.ObjInArgsAlg(0, 0, 0, 0, 0, 0, 0, 0) if kind == "oo" else ia
        ret, w1 = one.process_batch(nxt, follow)
        for c in codecs:
            ret, g1 = c.process_batch(nxt, follow)
            assert ret == 0 and bytes(memoryview(g1[0]))[:nb] == bytes(memoryview(w1[0]))[:nb], kind
    for c in codecs + [one]:
        c.close()


def test_multi_rejects_what_it_cannot_shard():
    w, h = 320, 240
    a, b = open_sensor("wl", w, h), open_sensor("wl", 160, 120)
    frames = synth.make_batch("scene", range(4), w, h, "yuyv")
    ia = xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 40, 0)
    ret, _ = a.process_batch(frames, ia, multi=[a, b])          # different geometry
    assert ret != 0 and "geometry" in sensors.last_error()
    ret, _ = a.process_batch(frames, ia, multi=[a, a])          # the same handle twice
    assert ret != 0
    c = open_sensor("wl", w, h)
    ret, _ = a.process_batch(frames, ia, multi=[a, c], flags=xdm.BATCH_ASYNC)
    assert ret != 0
    for x in (a, b, c):
        x.close()


def test_c_caller_shards_one_batch_over_the_gpus(tmp_path):
    """tests/c/multi_check.c: a plain C program, one process, one handle per visible GPU (at least two handles), through
    the C ABI only"""
    exe = str(tmp_path / "multi_check")
    subprocess.run(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "c", "multi_check.c"), "-ldl"], check=True)
    res = subprocess.run([exe, build.LIB], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "identical" in res.stdout
