#!/usr/bin/env python
"""bench.py -- frames/s of the sensor hot path on B200 (BASELINE.json metric), one JSON line.

Workload (config.workload): BASELINE.json configs[1], "webcam line sensor, batch of 4096 synthetic
320x240 YUYV frames on 1 B200".  One step = one pass of the hot path over the 4096-frame batch.

  value     frames/s with the batch already resident in HBM (device pointers into the C ABI,
            TRIKB200_BATCH_ASYNC, results into device memory), CUDA events on the launching stream.
            The batch (629 MB) is larger than L2 (126 MB), so no L2 flush is needed between steps.
  e2e       the same metric through the reference-facing call with HOST buffers:
            trikb200_processBatch(MEM_HOST) -- H2D of the frames from pinned memory, kernel, D2H of
            the OutArgs -- every step.
  roofline  algorithmic bytes (W*H*2 per frame) / kernel time, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the reference's own C/C++ (oracle/_ref, built for the host with the C6x emulation
            header) on a bounded sample of the same frames, on this box's host cores.

N > 1 (torchrun): frames shard by batch across ranks, no collective on the data path ("weak":
every rank runs the full per-GPU batch); barrier + max-over-ranks timing through torch.distributed.

--impl reference: the reference arm -- the host-built reference on all host cores (rank 0 only).
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 320, 240
BATCH = 4096
UNIQUE = 512                      # distinct synthetic frames, tiled to BATCH distinct device addresses
IN_ARGS = (0, 359, 0, 100, 0, 40, 0)   # detectValFrom=0, detectValTo=40
KIND = "wl"
METRIC = "frames_per_sec"
UNIT = "frames/s"
WORKLOAD = "webcam line sensor (WL), batch of %d synthetic %dx%d YUYV frames per GPU" % (BATCH, W, H)


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this same command (profiles/ncu_wl_summary.json)."""
    path = os.path.join(ROOT, "profiles", "ncu_wl_summary.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("dram_bytes_per_launch"), d.get("source")
    return None, None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None
        self.thread = None
        self.split = None

    def mark(self):
        """Samples from here on belong to the sustained phase."""
        self.split = len(self.samples)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        def summarise(samples):
            sm, mx, pw, reasons = [], [], [], set()
            for s in samples:
                parts = [p.strip() for p in s.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                    pw.append(float(parts[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                    "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}
        split = len(self.samples) if self.split is None else self.split
        out = summarise(self.samples[:split])
        out["phase"] = "warm-up + timed K steps (x repeats)"
        if self.split is not None:
            out["sustained_phase"] = summarise(self.samples[split:])
        return out


# ---------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, seeds, reps = args
    from oracle import ref as oracle
    from trik_media_sensors_dsp_b200 import synth
    frames = [synth.make_frame("scene", s, W, H, "yuyv") for s in seeds]
    aligned = []
    for f in frames:
        a = oracle.aligned_bytes(f.size)
        a[:] = f
        aligned.append(a)
    use_ref = oracle.ref_available(kind)
    if use_ref:
        sensor = oracle.RefSensor(kind)
        code, _ = sensor.setup(W, H)
        assert code == 0
    else:
        sensor = oracle.OracleSensor(kind, W, H)
    ia = oracle.RangeInArgs(*IN_ARGS)
    sensor.process(aligned[0], ia)           # warm-up
    t0 = time.perf_counter()
    n = 0
    for _ in range(reps):
        for a in aligned:
            sensor.process(a, ia)
            n += 1
    return n, time.perf_counter() - t0, use_ref


def cpu_frames_per_sec(cores, frames_per_core, reps=1):
    seeds = [list(range(c * frames_per_core, (c + 1) * frames_per_core)) for c in range(cores)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker((KIND, seeds[0], reps))]
    else:
        with ctx.Pool(cores) as pool:
            res = pool.map(_cpu_worker, [(KIND, s, reps) for s in seeds])
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return total / slowest, total, res[0][2], wall


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = 96
    vals = []
    for _ in range(args.warmup):
        cpu_frames_per_sec(cores, 16)
    t0 = time.perf_counter()
    frames = 0
    for _ in range(args.steps):
        fps, n, is_ref, _ = cpu_frames_per_sec(cores, per_core)
        vals.append(fps)
        frames += n
    wall = time.perf_counter() - t0
    value = float(np.median(vals))
    sample = "%d frames per step (%d per core) of the %d-frame workload, one process per core" % (per_core * cores, per_core, BATCH)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * BATCH / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int16 lanes", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sensor": KIND, "width": W, "height": H, "batch_per_gpu": BATCH},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if is_ref else "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sustained", action="store_true", help="skip the 1.5 s steady-state phase")
    ap.add_argument("--no-repeats", action="store_true", help="time the K steps once only")
    ap.add_argument("--no-others", action="store_true", help="skip the short resident measurement of the other four sensors")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from trik_media_sensors_dsp_b200 import open_sensor, synth, xdm, launch_count, build

    build.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # one process per GPU, bound to the GPU's NUMA node before any pinned buffer exists (e2e leg at N > 1)
    from trik_media_sensors_dsp_b200 import sharding
    bound = sharding.bind_to_gpu_numa(local_rank)

    n = args.batch
    fbytes = synth.frame_bytes(W, H, "yuyv")
    # this rank's shard of the stream: seeds rank*n .. rank*n + n - 1 (UNIQUE distinct frames, tiled)
    uniq = min(UNIQUE, n)
    host_unique = synth.make_batch("scene", [rank * n + i for i in range(uniq)], W, H, "yuyv")
    host = torch.empty((n, fbytes), dtype=torch.uint8, pin_memory=True)
    hv = host.numpy()
    for i in range(0, n, uniq):
        hv[i:i + uniq] = host_unique[:min(uniq, n - i)]
    d_frames = host.to(dev)
    rec = C.sizeof(xdm.TargetOutArgsAlg)
    d_out = torch.zeros((n, rec), dtype=torch.uint8, device=dev)
    h_out = torch.zeros((n, rec), dtype=torch.uint8, pin_memory=True)

    from trik_media_sensors_dsp_b200 import lib
    lib().trikb200_setDevice(local_rank)
    codec = open_sensor(KIND, W, H)
    ia = xdm.RangeInArgsAlg(*IN_ARGS)
    stream = torch.cuda.Stream(dev)              # a real (non-NULL) stream: NULL would mean "the handle's own stream"
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0

    def step_resident():
        ret, _ = codec.process_batch(d_frames.data_ptr(), ia, frames_device=True, frame_stride=fbytes, num_frames=n,
                                     out_device_ptr=d_out.data_ptr(), stream=sptr, flags=xdm.BATCH_ASYNC)
        assert ret == 0

    out_arr = (xdm.TargetOutArgsAlg * n).from_buffer(h_out.numpy())

    def step_e2e():
        ret, _ = codec.process_batch(hv, ia, out_algs=out_arr, stream=sptr)
        assert ret == 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t.item())

    # correctness guard: the resident path and the host path agree on this batch
    step_resident()
    torch.cuda.synchronize(dev)
    step_e2e()
    assert bytes(d_out.cpu().numpy()[:, :3].tobytes()) == bytes(h_out.numpy()[:, :3].tobytes()), "resident vs host path mismatch"

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.4)                       # let nvidia-smi start printing before the GPU gets busy
    # Phase 1 -- the contract: W untimed warm-up steps, then EXACTLY K timed steps.  Repeated a few times
    # (each repeat is W + K steps) only so that the 100 ms nvidia-smi sampler has samples that fall inside
    # timed regions; the reported value is the FIRST repeat's K steps, the others are listed beside it.
    repeats = 1 if args.no_repeats else 8
    runs = []
    l0 = launch_count()
    for rep in range(repeats):
        for _ in range(args.warmup):
            step_resident()
        if rep == 0:
            l0 = launch_count()
        runs.append(timed(step_resident, args.steps))
        if rep == 0:
            launches = launch_count() - l0
    ms_total = runs[0]
    t_burst_end = time.time()
    # Phase 2 -- the same step back to back for ~1.5 s: the steady state a streaming deployment sees
    # (on a 1 kW part this memory-bound kernel reaches the software power cap and the SM clock drops).
    sustained = None
    if not args.no_sustained:
        sampler.mark()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        t_end = time.perf_counter() + 1.5
        n_sus = 0
        while time.perf_counter() < t_end:
            for _ in range(50):
                step_resident()
            n_sus += 50
            torch.cuda.synchronize(dev)
        s1.record(stream)
        torch.cuda.synchronize(dev)
        sus_ms = s0.elapsed_time(s1) / n_sus
        sustained = {"steps": n_sus, "ms_per_step": sus_ms, "value": world * n / (sus_ms / 1000.0)}
    time.sleep(0.15)
    clocks = sampler.stop() if rank == 0 else None

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, e2e_steps)

    ms_per_step = ms_total / args.steps
    value = world * n / (ms_per_step / 1000.0)
    e2e_value = world * n / (ms_e2e / e2e_steps / 1000.0)

    if rank == 0:
        peak, peak_src = peaks()
        traffic, traffic_src = ncu_traffic()
        algo_bytes = n * W * H * 2                     # per launch (one launch per step per GPU)
        achieved = algo_bytes / (ms_per_step / 1000.0) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int16 lanes", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sensor": KIND, "width": W, "height": H, "batch_per_gpu": n,
                       "in_args": list(IN_ARGS), "l2": "inputs (%.0f MB per GPU) larger than L2, no flush" % (n * fbytes / 1e6),
                       "parallelism": "frames sharded by batch across %d GPU(s), no data-path collective" % world},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic if n == BATCH else None, "traffic_source": traffic_src,
                         "peak_source": peak_src, "kernel": "vsum_kernel<YUYV> (WL)",
                         "algorithmic_bytes_per_launch": algo_bytes},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * fbytes, "d2h_bytes_per_step": n * rec,
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "repeat_ms_per_step": [r / args.steps for r in runs],
            "sustained": sustained,
        }
        line["config"]["host_affinity"] = ("rank bound to its GPU's NUMA node: %d of %d CPUs" % (len(bound[1]), len(bound[0]))
                                           if bound else "unbound (NVML affinity query unavailable)")
        if not args.no_others and world == 1:
            # beside the headline: the other four sensors on the same batch shape, frames resident, 10 steps each
            # (explanatory only -- `value` above is the BASELINE metric)
            others = {}
            from trik_media_sensors_dsp_b200 import sensors as _sensors
            for kind in ("ol", "wo", "om", "oo"):
                layout = _sensors.layout_of(xdm.KIND_OF[kind])
                fam = "grid" if kind == "om" else "scene"
                hu = synth.make_batch(fam, range(uniq), W, H, layout)
                hv[:] = 0
                for i in range(0, n, uniq):
                    hv[i:i + uniq] = hu[:min(uniq, n - i)]
                d_frames.copy_(host)
                k = xdm.KIND_OF[kind]
                orec = C.sizeof(xdm.OUT_ARGS_ALG[k])
                o_out = torch.zeros((n, orec), dtype=torch.uint8, device=dev)
                oc = open_sensor(kind, W, H)
                oia = (xdm.ObjInArgsAlg(1, 0, 20, 80, 20, 50, 30, 0) if kind == "oo"
                       else (xdm.MxnInArgsAlg(3, 3) if kind == "om" else xdm.RangeInArgsAlg(0, 359, 0, 100, 0, 40, 0)))

                def other_step():
                    r, _ = oc.process_batch(d_frames.data_ptr(), oia, frames_device=True, frame_stride=fbytes, num_frames=n,
                                            out_device_ptr=o_out.data_ptr(), stream=sptr, flags=xdm.BATCH_ASYNC)
                    assert r == 0

                for _ in range(3):
                    other_step()
                oms = timed(other_step, 10) / 10
                others[kind] = {"frames_per_sec": n / (oms / 1000.0), "ms_per_step": oms,
                                "hbm_frac": n * W * H * 2 / (oms / 1000.0) / 1e9 / peak, "frames": fam}
                oc.close()
                del o_out
            line["other_sensors"] = others
        if not args.no_cpu and world == 1:
            if bound:
                os.sched_setaffinity(0, bound[0])          # the CPU baseline uses every core of the box again
            cores = os.cpu_count() or 1
            fps1, n1, is_ref, _ = cpu_frames_per_sec(1, 128)
            fpsN, nN, _, _ = cpu_frames_per_sec(cores, 128)
            line["cpu_baseline"] = {"value": fpsN, "unit": UNIT, "cores": cores, "kind": "reference" if is_ref else "port",
                                    "sample": "%d frames (128 per core) of the same synthetic workload; single core: %.1f frames/s" % (nN, fps1),
                                    "single_core_value": fps1}
        print(json.dumps(line), flush=True)
    codec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
