"""Where an active-set change spends its time: cProfile of FBPINNTrainer.set_active at the full cfg 5 size
(or cfg 3 with --cfg3), once with asynchronous launches and once with CUDA_LAUNCH_BLOCKING=1 semantics emulated by a
synchronize after every library call.  Prints the top cumulative entries and writes gpurun_out/rebuild_profile.txt.
    python tests/tools/rebuild_profile.py [--cfg3]"""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbpinns_b200 import configs                    # noqa: E402
from fbpinns_b200.trainers import FBPINNTrainer     # noqa: E402
from fbpinns_b200.util.logger import logger         # noqa: E402

logger.setLevel("WARNING")
cfg3 = "--cfg3" in sys.argv
c = (configs.cfg3_burgers(device="cuda:0", line_scheduler=False) if cfg3 else configs.cfg5_poisson(device="cuda:0"))
tr = FBPINNTrainer(c)
tr.setup()
m = tr.dd.m
act = np.ones(m, dtype=int)
tr.set_active(act)
for _ in range(6):
    tr.step()
torch.cuda.synchronize()
out = io.StringIO()
ts = []
for rep in range(3):
    t0 = time.perf_counter()
    tr.set_active(act)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
    t1 = time.perf_counter()
    for _ in range(5):
        tr.step()
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t1) * 1e3)
print("set_active ms / first 5 steps (3 eager + capture + replay) ms, three repetitions:", [round(t, 2) for t in ts], file=out)
pr = cProfile.Profile()
pr.enable()
tr.set_active(act)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(35)
txt = out.getvalue()
print(txt)
d = os.path.join(ROOT, "gpurun_out")
os.makedirs(d, exist_ok=True)
open(os.path.join(d, "rebuild_profile_cfg3.txt" if cfg3 else "rebuild_profile.txt"), "w").write(txt)
