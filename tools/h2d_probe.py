"""Bare pinned-memory copy bandwidth, one process per GPU, all ranks copying at the same time:

    python tools/h2d_probe.py                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/h2d_probe.py

What the end-to-end numbers of bench.py (e2e: host buffers in, list out) are bounded by when N ranks share
the host: prints per-rank and aggregate H2D / D2H GB/s, the ranks starting together behind a barrier.
("concurrent N-rank bare cudaMemcpyAsync probe".)
"""
import json
import os

import torch

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {}
for name, dst, src in (("h2d", d, h), ("d2h", h2, d)):
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    out[name] = 8 * n / e0.elapsed_time(e1) / 1e6
if world > 1:
    t = torch.tensor([out["h2d"], out["d2h"]], device="cuda", dtype=torch.float64)
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    if rank == 0:
        per = [[round(float(x[0]), 1), round(float(x[1]), 1)] for x in g]
        print(json.dumps({"ranks": world, "h2d_gbs_per_rank": [p[0] for p in per], "d2h_gbs_per_rank": [p[1] for p in per],
                          "h2d_gbs_sum": round(sum(p[0] for p in per), 1), "d2h_gbs_sum": round(sum(p[1] for p in per), 1),
                          "bytes_per_copy": n, "host": "pinned (cudaHostAlloc via torch), all ranks at once"}))
    dist.destroy_process_group()
else:
    print(json.dumps({"ranks": 1, "h2d_gbs": round(out["h2d"], 1), "d2h_gbs": round(out["d2h"], 1), "bytes_per_copy": n}))
