#!/usr/bin/env python
"""BASELINE.json config 4: mixed line + object + mxn sensor instances over many concurrent synthetic
streams, streams sharded round-robin across the GPUs of one box.

    python tools/mixed_streams.py [--streams 1024] [--frames 8] [--check 40]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/mixed_streams.py --gpus N

Every time step feeds ONE frame of every stream of this rank through trikb200_processMixed (host frames,
one handle per stream, handles overlap on their own CUDA streams).  A sample of the results is re-computed
with sequential process() calls on fresh handles (parity against the oracle is the job of tests/); rank 0
prints one JSON line (whole-job frames/s, wall clock, max over ranks)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KINDS = ["wo", "wl", "ol", "oo", "om"]
W, H = 320, 240


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--streams", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--check", type=int, default=40)
    ap.add_argument("--mode", default="streams", choices=["streams", "handles"],
                    help="streams: one handle per sensor kind, every stream a logical stream of one batch "
                         "(TRIKB200_Batch.streamIds); handles: one codec handle per stream (trikb200_processMixed)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import torch
    import torch.distributed as dist
    from trik_media_sensors_dsp_b200 import open_sensor, process_mixed, sensors, sharding, synth, xdm, launch_count

    torch.cuda.set_device(local_rank)
    from trik_media_sensors_dsp_b200 import sharding as _sh
    _sh.bind_to_gpu_numa(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sensors.lib().trikb200_setDevice(local_rank)

    mine = sharding.streams_of_rank(args.streams, world, rank)

    def in_alg(kind, t):
        if kind == "oo":
            return xdm.ObjInArgsAlg(1, 0, 20, 80, 20, 50, 30, 0) if t == 0 else xdm.ObjInArgsAlg(0, 0, 0, 0, 0, 0, 0, 0)
        if kind == "om":
            return xdm.MxnInArgsAlg(3, 3)
        return xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 40, 0)

    # frames: 16 distinct per kind, reused round robin (generation is not what is measured)
    pool = {}
    for kind in KINDS:
        fam = "blobs" if kind == "oo" else ("grid" if kind == "om" else "scene")
        pool[kind] = [synth.make_frame(fam, i, W, H, sensors.layout_of(xdm.KIND_OF[kind])) for i in range(16)]

    results = {}                                                          # (stream, t) -> OutArgsAlg
    if args.mode == "handles":
        codecs = {s: open_sensor(KINDS[s % 5], W, H) for s in mine}
        steps = []
        for t in range(args.frames):
            items = []
            for s in mine:
                kind = KINDS[s % 5]
                oa = xdm.OUT_ARGS_ALG[xdm.KIND_OF[kind]]()
                results[(s, t)] = oa
                items.append((codecs[s], pool[kind][(s + t) % 16], in_alg(kind, t), oa, 7))
            steps.append(items)
        assert process_mixed(steps[0][:min(64, len(steps[0]))]) == 0      # warm-up (allocations, module load)
        for s in mine:                                                    # restart every stream's carried state
            assert codecs[s].set_params(W, H) == 0

        def run_step(t):
            assert process_mixed(steps[t]) == 0, sensors.last_error()
    else:
        codecs = {kind: open_sensor(kind, W, H) for kind in KINDS}
        by_kind = {kind: [s for s in mine if KINDS[s % 5] == kind] for kind in KINDS}
        plan = {}
        for t in range(args.frames):
            for kind in KINDS:
                ss = by_kind[kind]
                if not ss:
                    continue
                frames = np.stack([pool[kind][(s + t) % 16] for s in ss])
                InAlg, OutAlg = xdm.IN_ARGS_ALG[xdm.KIND_OF[kind]], xdm.OUT_ARGS_ALG[xdm.KIND_OF[kind]]
                ias = (InAlg * len(ss))(*[in_alg(kind, t) for _ in ss])
                outs = (OutAlg * len(ss))()
                for j, s in enumerate(ss):
                    results[(s, t)] = outs[j]
                plan[(t, kind)] = (frames, ias, outs, list(range(len(ss))), len(ss))
        for kind in KINDS:                                                # warm-up, then restart the carried state
            if by_kind[kind]:
                frames, ias, outs, ids, ns = plan[(0, kind)]
                assert codecs[kind].process_batch(frames, ias, stream_ids=ids, num_streams=ns)[0] == 0
                assert codecs[kind].set_params(W, H) == 0

        def run_step(t):
            for kind in KINDS:
                if by_kind[kind]:
                    frames, ias, outs, ids, ns = plan[(t, kind)]
                    ret, _ = codecs[kind].process_batch(frames, ias, out_algs=outs, stream_ids=ids, num_streams=ns, seeds=[7] * ns)
                    assert ret == 0, sensors.last_error()

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = launch_count()
    t0 = time.perf_counter()
    for t in range(args.frames):
        run_step(t)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    launches = launch_count() - l0
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())

    # self-consistency sample: the same frames through sequential process() calls on a fresh handle per stream
    bad = 0
    for s in mine[:args.check]:
        kind = KINDS[s % 5]
        fresh = open_sensor(kind, W, H)
        for t in range(args.frames):
            ret, oa = fresh.process(pool[kind][(s + t) % 16], in_alg(kind, t), seed=7)
            got = results[(s, t)]
            n = {"om": 36, "oo": 24}.get(kind, 3)
            if ret != 0 or bytes(memoryview(got))[:n] != bytes(memoryview(oa.alg))[:n]:
                bad += 1
        fresh.close()
    tb = torch.tensor([bad], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tb)
    if rank == 0:
        total = args.streams * args.frames
        print(json.dumps({"config": "mixed WO/WL/OL/OO/OM instances, %d streams x %d frames of %dx%d, streams round-robin over %d GPU(s)"
                          % (args.streams, args.frames, W, H, world), "frames_per_sec": total / dt, "wall_s": dt,
                          "n_gpus": world, "gpu_launches_rank0": int(launches), "mismatches_vs_sequential_process": int(tb.item()),
                          "mode": args.mode,
                          "path": ("one handle per sensor kind, TRIKB200_Batch.streamIds" if args.mode == "streams" else "trikb200_processMixed, one handle per stream")
                                  + ", host frames (H2D inside the timed region)"}), flush=True)
    for c in codecs.values():
        c.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
