"""Kernel list of ONE training step (every CUDA kernel with its duration, in launch order) from torch.profiler, for the
single-GPU step or, under torchrun, for rank 0 of the sharded step.  Shows what the step consists of besides the two
subdomain kernels (the fixed per-step latency that limits strong scaling).
    python tests/tools/step_kernel_list.py [--shape a,b,c,d]          torchrun ... tests/tools/step_kernel_list.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbpinns_b200 import configs                     # noqa: E402
from fbpinns_b200.trainers import FBPINNTrainer      # noqa: E402
from fbpinns_b200.util.logger import logger          # noqa: E402

logger.setLevel("WARNING")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
kw = {}
if "--shape" in sys.argv:
    a, b, c_, d = (int(v) for v in sys.argv[sys.argv.index("--shape") + 1].split(","))
    kw = dict(n_sub=(a, b), n_pts=(c_, d))
graph = "--graph" in sys.argv        # profile a CUDA-graph replay (what the bench times) instead of eager launches
c = configs.cfg5_poisson(device=f"cuda:{local}", use_cuda_graph=graph, **kw)
tr = FBPINNTrainer(c)
if world > 1:
    from fbpinns_b200.parallel import shard_trainer
    shard_trainer(tr, rank, world)
tr.setup()
tr.set_active(np.ones(tr.dd.m, dtype=int))
for _ in range(5):
    tr.step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
from torch.profiler import profile, ProfilerActivity   # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step()
    torch.cuda.synchronize()
show = int(os.environ.get("LIST_RANK", 0))
if rank == show:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    tot = 0.0
    lines = []
    prev_end = None
    for e in evs:
        dur = e.time_range.elapsed_us()
        tot += dur
        gap = (e.time_range.start - prev_end) if prev_end is not None else 0.0
        prev_end = max(prev_end or 0, e.time_range.end)
        lines.append(f"{dur:9.1f} us  (+{gap:6.1f} gap)  {e.name[:100]}")
    span = (evs[-1].time_range.end - evs[0].time_range.start) if evs else 0
    txt = "\n".join(lines) + f"\n--- {len(evs)} kernels, sum {tot:.1f} us, first-to-last span {span:.1f} us (eager launch, world {world})\n"
    print(txt)
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    open(os.path.join(d, f"step_kernels_n{world}.txt"), "w").write(txt)
if world > 1:
    dist.barrier()
    os._exit(0)
