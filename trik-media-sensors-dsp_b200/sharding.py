"""Multi-GPU plumbing: frames (or whole streams) shard across the ranks of one box with no
collective on the pixel path; the only exchange is the final gather of the OutArgs records
(<= 400 bytes per frame), SURVEY.md section 8(e).

One process per GPU (torchrun); torch.distributed is used for the rendezvous, the barrier and the
result gather only ("nccl" on GPUs, "gloo" in the CPU tests).
"""
import os

import numpy as np


def bind_to_gpu_numa(cuda_device_index):
    """Pin this process (thread) to the CPUs next to its GPU, so that the pinned staging buffers it allocates
    afterwards and the host side of its H2D copies stay on the GPU's NUMA node.  With eight ranks on a
    two-socket box the unbound default sends about half of the host traffic across the socket link.
    Returns (previous affinity, new affinity), or None when NVML or the device cannot be queried -- binding is
    a placement hint, never a requirement."""
    try:
        import pynvml
        import torch
        before = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(cuda_device_index).uuid)
        if not uuid.startswith("GPU-"):
            uuid = "GPU-" + uuid
        handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return before, os.sched_getaffinity(0)
    except Exception:
        return None


def partition(n, world, rank):
    """Contiguous, balanced slice [lo, hi) of n items for this rank (first n % world ranks get one more)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def streams_of_rank(num_streams, world, rank):
    """Whole streams (codec handles) assigned round-robin, so per-handle state never crosses a GPU."""
    return list(range(rank, num_streams, world))


def gather_records(local, n_total, world, rank, device=None):
    """Gather per-rank (n_local, record_bytes) uint8 arrays, partitioned by `partition`, to every rank
    in frame order.  Returns an (n_total, record_bytes) uint8 numpy array."""
    if world == 1:
        return np.ascontiguousarray(local)
    import torch
    import torch.distributed as dist
    rec = local.shape[1]
    counts = [partition(n_total, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in counts)
    pad = np.zeros((width, rec), dtype=np.uint8)
    pad[:local.shape[0]] = local
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full = np.empty((n_total, rec), dtype=np.uint8)
    for r, (lo, hi) in enumerate(counts):
        full[lo:hi] = outs[r].cpu().numpy()[:hi - lo]
    return full
