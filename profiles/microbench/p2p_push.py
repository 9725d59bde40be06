"""Delivery of a rank's rows to the other ranks' tables: per-peer copies through CUDA IPC mappings against
all_gather_into_tensor (torchrun --nproc-per-node N profiles/microbench/p2p_push.py)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist

import imfeat_b200 as imf

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
n_total, width = 100000 * world, 720
for transport in ("p2p", "collective"):
    tab = imf.distributed.ShardedTable(n_total, width, device=dev, transport=transport)
    per = tab.per
    tab.local_rows(0, per).fill_(float(rank + 1))
    for slab in (per, 16384):
        for rep in range(3):
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            for a, b in imf.distributed.slab_bounds(per, slab):
                tab.push(a, b)
            tab.side.synchronize()
            t1 = time.perf_counter()
            tab.finish()
            t2 = time.perf_counter()
        ok = all(float(tab.full[r * per].mean().item()) == r + 1 for r in range(world))
        if rank == 0:
            mb = per * width * 8 / 1e6
            print("%-10s (%s) slab %6d: push+sync %.2f ms  (%.0f GB/s out per rank to %d peers), finish %.2f ms, ok=%s" % (
                transport, tab.transport, slab, 1e3 * (t1 - t0), mb * (world - 1) / 1e3 / (t1 - t0), world - 1, 1e3 * (t2 - t1), ok), flush=True)
    del tab
    torch.cuda.empty_cache()
dist.destroy_process_group()
