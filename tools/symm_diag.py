#!/usr/bin/env python
"""2-GPU diagnostic of the symmetric-memory gather pieces: device time of barrier / local copy /
peer copy on a side stream, idle and while the persistent transform kernel runs on the main stream."""
import os, sys, time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nnaudio_b200 as nb  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mod = nb.features.MelSpectrogram(sr=22050, n_fft=2048, hop_length=512, n_mels=128, verbose=False).to(dev)
x = torch.randn(64, 220500, device=dev)
with torch.no_grad():
    y = mod(x)
shape = (2, world * y.shape[0]) + tuple(y.shape[1:])
buf = symm_mem.empty(shape, dtype=y.dtype, device=dev)
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
peers = [hdl.get_buffer(r, shape, y.dtype) for r in range(world)]
s = torch.cuda.Stream(dev)
n = y.shape[0]
peer = (rank + 1) % world


def timed(fn, busy, reps=20):
    """average device time of fn() on stream s; busy: keep the main stream running the transform"""
    torch.cuda.synchronize(); dist.barrier()
    evs = []
    with torch.no_grad():
        for i in range(reps):
            if busy:
                mod(x)
            with torch.cuda.stream(s):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(s); fn(); b.record(s)
            evs.append((a, b))
        if busy:
            mod(x)
    t0 = time.perf_counter()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs[2:]) / (reps - 2)


res = {}
for busy in (False, True):
    k = "busy" if busy else "idle"
    res[f"barrier_{k}"] = timed(lambda: hdl.barrier(channel=0), busy)
    res[f"local_copy_{k}"] = timed(lambda: peers[rank][0, rank * n:(rank + 1) * n].copy_(y, non_blocking=True), busy)
    res[f"peer_copy_{k}"] = timed(lambda: peers[peer][0, rank * n:(rank + 1) * n].copy_(y, non_blocking=True), busy)
    res[f"peer_copy_from_{k}"] = timed(lambda: buf[1, peer * n:(peer + 1) * n].copy_(peers[peer][0, peer * n:(peer + 1) * n], non_blocking=True), busy)

# host-side cost of enqueueing one gather
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for i in range(50):
    with torch.cuda.stream(s):
        hdl.barrier(channel=0)
        peers[peer][0, rank * n:(rank + 1) * n].copy_(y, non_blocking=True)
        hdl.barrier(channel=1)
res["host_enqueue_us_per_gather"] = (time.perf_counter() - t0) / 50 * 1e6
torch.cuda.synchronize()
with torch.no_grad():
    t0 = time.perf_counter()
    for i in range(50):
        mod(x)
    res["host_enqueue_us_per_transform"] = (time.perf_counter() - t0) / 50 * 1e6
torch.cuda.synchronize()
if rank == 0:
    mb = y.numel() * 4 / 1e6
    print(f"shard {mb:.1f} MB")
    for k, v in res.items():
        extra = f"  ({mb / v:.0f} GB/s)" if "copy" in k else ""
        print(f"{k:34s} {v:9.3f} {'us' if 'host' in k else 'ms'}{extra}")
dist.barrier(); dist.destroy_process_group()
