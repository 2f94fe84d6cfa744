"""Phase timeline of tc_backward_kernel2 (FBP_TC_DEBUG=64): thread 0 of block 0 stamps the cycle counter at the phase boundaries of
every tile of its work item.  Prints the average cycles per phase.  Needs a library built with the trace code:
    make -C fbpinns_b200/csrc BUILD=build_trace LIB=libfbpinn_b200_trace.so EXTRA=-DFBP_B2_TRACE
    FBP_LIB=$PWD/fbpinns_b200/csrc/libfbpinn_b200_trace.so python tests/tools/bwd_phase_trace.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbpinns_b200 import configs, _lib                                   # noqa: E402
from fbpinns_b200._lib import ptr, stream_ptr, check                     # noqa: E402
from fbpinns_b200.engine import Plan, ConstraintEvaluator                # noqa: E402
from fbpinns_b200.trainers import FBPINNTrainer                          # noqa: E402
from fbpinns_b200.util.logger import logger                              # noqa: E402

logger.setLevel("WARNING")
lib = _lib.load()
tr = FBPINNTrainer(configs.cfg5_poisson(device="cuda:0", kernel="tensor-full", use_cuda_graph=False)).setup()
tr.set_active(np.ones(tr.dd.m, dtype=int))
ev = tr.inputs.evaluators[0]
torch.manual_seed(0)
ubar = torch.randn(ev.takes.n, ev.V, device="cuda")
g = torch.zeros((ev.takes.m_active, tr.params.shape[1]), device="cuda")
ev.forward(tr.params)
ev.backward(ubar, tr.params, g, accumulate=False)
torch.cuda.synchronize()
dbgbuf = torch.zeros(64 * 8 + 64 * 4, device="cuda")
TID = int(sys.argv[1]) if len(sys.argv) > 1 else 0          # the recording thread: 0 = quarter 0, 64 = quarter 2 (unit group 0)
os.environ["FBP_TC_DEBUG"] = str(64 + 256 * TID)
tv = ev.takes.view()
for _ in range(3):
    check(lib.fbp_backward(ev.plan.handle, C.byref(tv), ptr(ev.x), ptr(tr.params), ptr(tr.dd.sub_static), ptr(ev.grow), ptr(g), 0,
                           ptr(ev.gpart), ptr(ev.scratch), ev.scratch_floats, ptr(dbgbuf), stream_ptr()), "fbp_backward")
torch.cuda.synchronize()
raw = dbgbuf.cpu().numpy()
t = raw[:512].reshape(64, 8)
rs = raw[512:].reshape(64, 4)
nt = int((t[:, 7] > 0).sum())
t = t[:nt]
names = ["L (-> a1 arrive)", "wait MMA1 (+gather)", "E1 (-> a3 arrive)", "ring store (q<2)", "wait MMA3", "E2", "S0 of the next tile + ring store (q>=2)"]
print(f"thread {TID}: tiles of block 0's item: {nt}; cycles per tile (mean over tiles 1..): {np.diff(t[:, 0])[0:].mean():.0f}")
for k, nm in enumerate(names):
    d = t[1:, k + 1] - t[1:, k]
    print(f"  {nm:24s} {d.mean():8.0f}  (min {d.min():.0f}, max {d.max():.0f})")
print("  tile boundary (end -> next start):", (t[1:, 0] - t[:-1, 7]).mean())

rs = rs[1:nt]
print("ring store of warp 0 (cycles): slot wait -> %0.f, TMEM reload %.0f, splits + 128 STS %.0f, proxy fence + arrive %.0f" % (
    (rs[:, 0] - t[1:nt, 3]).mean(), (rs[:, 1] - rs[:, 0]).mean(), (rs[:, 2] - rs[:, 1]).mean(), (rs[:, 3] - rs[:, 2]).mean()))
if TID & 64:     # quarters 2, 3: the store follows E2 and S0 of the next tile
    print("late store: E2 done -> slot acquired (S0 of the next tile + slot wait) %.0f" % (rs[:, 0] - t[1:nt, 6]).mean())
