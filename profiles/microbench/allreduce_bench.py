"""All-reduce of the dense gradient buffer (59 rows x 1M Gaussians = 236 MB fp32): ncclAllReduce vs the own NVLS kernel
(csrc/nvls.cu), CUDA events, max over ranks.   torchrun --nproc-per-node N profiles/microbench/allreduce_bench.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from litegs_b200 import dist as lgs_dist  # noqa: E402
from litegs_b200.dist import PARAM_ORDER  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
C, S = 7813, 128
P = {"xyz": torch.empty((3, C, S), device=dev), "scale": torch.empty((3, C, S), device=dev), "rot": torch.empty((4, C, S), device=dev),
     "sh_0": torch.empty((1, 3, C, S), device=dev), "sh_rest": torch.empty((15, 3, C, S), device=dev), "opacity": torch.empty((1, C, S), device=dev)}
out = {"world": world, "bytes": None}
for name, sym in (("nccl", False), ("own_nvls", True)):
    for ctas in ((0,) if not sym else (32, 64, 128, 256)):
        os.environ["LGS_NVLS_CTAS"] = str(ctas)
        acc = lgs_dist.GradAccumulator(P, symmetric=sym)
        if sym and acc._nvls is None:
            out[name] = "unavailable"
            break
        out["bytes"] = acc.flat_all.numel() * 4
        acc.flat_all.fill_(1.0)
        for _ in range(3):
            acc.all_reduce()
        torch.cuda.synchronize(); dist.barrier()
        ok = bool(torch.all(acc.flat_all == float(world) ** 3))
        acc.flat_all.fill_(1e-3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); dist.barrier()
        e0.record()
        for _ in range(10):
            acc.all_reduce()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[f"{name}{'_ctas' + str(ctas) if sym else ''}"] = {"ms": float(t), "exact": ok}
        del acc
if rank == 0:
    print("ALLREDUCE_BENCH " + json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
