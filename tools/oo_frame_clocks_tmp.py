import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
from trik_media_sensors_dsp_b200 import open_sensor, synth, xdm, sensors
w, h = int(sys.argv[1]), int(sys.argv[2]); fam = sys.argv[3]; wide = sys.argv[4] == "wide"
n, uniq = 4096, 64
layout = "yuv422p"
fbytes = synth.frame_bytes(w, h, layout)
hu = synth.make_batch(fam, range(uniq), w, h, layout)
host = np.concatenate([hu] * (n // uniq))
d_frames = torch.from_numpy(host).cuda()
rec = C.sizeof(xdm.OUT_ARGS_ALG[xdm.KIND_OF["oo"]])
d_out = torch.zeros((n, rec), dtype=torch.uint8, device="cuda")
codec = open_sensor("oo", w, h)
ia = xdm.ObjInArgsAlg(1, 0, 25, 75, 25, 60, 40, 0) if wide else xdm.ObjInArgsAlg(1, 0, 20, 80, 20, 50, 30, 0)
stream = torch.cuda.Stream(); sptr = C.c_void_p(stream.cuda_stream)
for _ in range(4):
    ret, _ = codec.process_batch(d_frames.data_ptr(), ia, frames_device=True, frame_stride=fbytes, num_frames=n,
                                 out_device_ptr=d_out.data_ptr(), stream=sptr, flags=xdm.BATCH_ASYNC)
    assert ret == 0
stream.synchronize()
o = d_out.cpu().numpy()
f = o[:, 24:36].copy().view(np.uint32).reshape(n, 3).astype(np.int64)
walk, total, t0 = f[:, 0], f[:, 1], f[:, 2]
t0 = (t0 - t0.min())
end_ns = t0 + total / 1.965
print(json.dumps({"size": [w, h], "family": fam, "wide": wide,
  "walk_cycles": {"mean": float(walk.mean()), "median": float(np.median(walk)), "p90": float(np.percentile(walk, 90)), "max": int(walk.max())},
  "total_cycles": {"mean": float(total.mean()), "max": int(total.max())},
  "start_ns_after_first": {"median": float(np.median(t0)), "p90": float(np.percentile(t0, 90)), "max": int(t0.max())},
  "kernel_span_us": float(end_ns.max() / 1000.0),
  "per_distinct_frame_walk_kcycles": [int(walk[i::uniq].mean() / 1000) for i in range(uniq)]}))
