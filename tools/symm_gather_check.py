#!/usr/bin/env python
"""2+-GPU check of the copy-engine gather (torchrun --nproc-per-node N tools/symm_gather_check.py):
the gathered tensor equals NCCL's all_gather for both modes, over several steps with slot reuse."""
import os, sys
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nnaudio_b200.parallel import BatchShardedTransform  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import nnaudio_b200 as nb  # noqa: E402

ok = True
# ---- peer gather (kernels write into symmetric memory, CE pushes, stream-memop handshakes) on a real module
mod = nb.features.STFT(n_fft=512, hop_length=128, output_format="Magnitude", verbose=False).to(dev)
trp = BatchShardedTransform(lambda t: mod(t), gather=True, reserve_sms=0)
for mode in ("all", "root"):
    for step in range(7):
        x = torch.randn(4, 16000, generator=torch.Generator(device=dev).manual_seed(100 * rank + step), device=dev)
        with torch.no_grad():
            work, got = trp.forward_async_peer(x, slot=step & 1, gather_to=mode)
            work.wait()
            mine = mod(x)
        want = torch.empty((world * 4,) + tuple(mine.shape[1:]), device=dev)
        dist.all_gather_into_tensor(want, mine)
        torch.cuda.synchronize()
        if mode == "all" or rank == 0:
            good = torch.equal(got, want)
            ok &= good
            if not good:
                print(f"rank {rank} peer mode {mode} step {step}: MISMATCH {(got - want).abs().max().item()}")
        trp.release_peer(step & 1, None if mode == "all" else [0])
    torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    print("peer gather:", "OK" if ok else "FAILED")

tr = BatchShardedTransform(lambda t: t * 2.0 + 1.0, gather=True, reserve_sms=0)
for mode in ("all", "root"):
    for step in range(6):
        x = torch.full((4, 128, 431), float(rank * 100 + step), device=dev) + torch.arange(431, device=dev)
        work, got = tr.forward_async_symm(x, slot=step & 1, gather_to=mode)
        work.wait()
        want = torch.empty((world * 4, 128, 431), device=dev)
        dist.all_gather_into_tensor(want, x * 2.0 + 1.0)
        torch.cuda.synchronize()
        if mode == "all" or rank == 0:
            good = torch.equal(got, want)
            ok &= good
            if not good:
                print(f"rank {rank} mode {mode} step {step}: MISMATCH {(got - want).abs().max().item()}")
        tr.release(step & 1)
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("symm gather check:", "OK" if int(flag.item()) else "FAILED", f"(world {world})")
dist.barrier()
dist.destroy_process_group()
