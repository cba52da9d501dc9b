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
tr = BatchShardedTransform(lambda t: t * 2.0 + 1.0, gather=True, reserve_sms=0)
ok = True
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
