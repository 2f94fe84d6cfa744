"""Launch time of the two tensor kernels at the full cfg 5 size (CUDA events, median of 20 launches each, L2 flushed in between)
for the library FBP_LIB points at: A/B runs of differently built libraries in one gpurun call.
    FBP_LIB=... python tests/tools/time_tc_kernels.py [tag]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbpinns_b200 import configs                                         # noqa: E402
from fbpinns_b200.trainers import FBPINNTrainer                          # noqa: E402
from fbpinns_b200.util.logger import logger                              # noqa: E402

logger.setLevel("WARNING")
tr = FBPINNTrainer(configs.cfg5_poisson(device="cuda:0", kernel="tensor-full", use_cuda_graph=False)).setup()
tr.set_active(np.ones(tr.dd.m, dtype=int))
ev = tr.inputs.evaluators[0]
torch.manual_seed(0)
ubar = torch.randn(ev.takes.n, ev.V, device="cuda")
g = torch.zeros((ev.takes.m_active, tr.params.shape[1]), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=20):
    ts = []
    for _ in range(reps + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts[3:]))


ev.forward(tr.params)
ev.backward(ubar, tr.params, g, accumulate=False)
f = timed(lambda: ev.forward(tr.params))
b = timed(lambda: ev.backward(ubar, tr.params, g, accumulate=False))
print(f"{sys.argv[1] if len(sys.argv) > 1 else os.environ.get('FBP_LIB', 'product')}: forward path {f:.3f} ms, reverse path {b:.3f} ms "
      f"(evaluator calls: subdomain kernel + reduce kernels), grad checksum {float(g.double().abs().sum()):.6e}")
