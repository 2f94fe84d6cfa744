"""Where the per-work-item cost of the tensor kernels goes (cfg 5 full size, 1 GPU): every CTA (= one work item) records its SM,
its entry / prologue-done / tile-loop-done / all-warps-done / exit cycle counts (fbp_tc.cuh blk_stamp).  Per SM the CTAs are
ordered by entry time; printed: mean prologue, tile loop (and per tile), drain (loop end of warp 0 -> every warp done),
end-of-item reduction, and the gap between one CTA's exit and the next CTA's entry on the same SM.  Needs a trace build:
    make -C fbpinns_b200/csrc BUILD=build_btrace LIB=libfbpinn_b200_btrace.so EXTRA=-DFBP_BLOCK_TRACE
    FBP_LIB=$PWD/fbpinns_b200/csrc/libfbpinn_b200_btrace.so python tests/tools/item_cost_trace.py"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbpinns_b200 import configs, _lib                                   # noqa: E402
from fbpinns_b200.trainers import FBPINNTrainer                          # noqa: E402
from fbpinns_b200.util.logger import logger                              # noqa: E402

logger.setLevel("WARNING")
lib = _lib.load()
set_trace = C.CDLL(_lib.LIB_PATH).fbp_debug_set_block_trace
set_trace.argtypes = [C.c_void_p]
kw = {} if "--small" not in sys.argv else dict(n_sub=(16, 32), n_pts=(362, 724))      # one rank's share at 8 GPUs
tr = FBPINNTrainer(configs.cfg5_poisson(device="cuda:0", kernel="tensor-full", use_cuda_graph=False, **kw)).setup()
tr.set_active(np.ones(tr.dd.m, dtype=int))
ev = tr.inputs.evaluators[0]
torch.manual_seed(0)
ubar = torch.randn(ev.takes.n, ev.V, device="cuda")
g = torch.zeros((ev.takes.m_active, tr.params.shape[1]), device="cuda")
n_items = ev.takes.n_items_active
res = {"items": int(n_items), "pairs": int(ev.takes.s)}


def analyse(buf, name):
    b = buf.cpu().numpy().view(np.uint32).reshape(-1, 8)[:n_items].astype(np.int64)
    sm, t = b[:, 0], b[:, 1:6]
    d = lambda a, c: (t[:, a] - t[:, c]) & 0xffffffff
    pro, loop, drain, red = d(1, 0), d(2, 1), d(3, 2), d(4, 3)
    nt = b[:, 6]
    gaps = []
    for s in np.unique(sm):
        idx = np.nonzero(sm == s)[0]
        idx = idx[np.argsort(b[idx, 7])]
        # entry of the next CTA on this SM minus exit of the previous one (same SM counter)
        gaps += [int((t[j, 0] - t[i, 4]) & 0xffffffff) for i, j in zip(idx[:-1], idx[1:])]
    gaps = np.array([x for x in gaps if x < 1 << 30])
    full = nt == np.bincount(nt).argmax()
    out = {"tiles_mode": int(np.bincount(nt).argmax()), "prologue": float(pro.mean()), "tile_loop": float(loop[full].mean()),
           "per_tile": float((loop[full] / nt[full]).mean()), "drain": float(drain.mean()), "reduction_and_exit": float(red.mean()),
           "gap_to_next_cta": float(np.median(gaps)), "gap_mean": float(gaps.mean()), "ctas_per_sm": float(len(b) / len(np.unique(sm)))}
    pt = loop[full] / nt[full]
    out["per_tile_p5_p50_p95"] = [float(np.percentile(pt, q)) for q in (5, 50, 95)]
    sm_mean = np.array([pt[sm[full] == s_].mean() for s_ in np.unique(sm[full])])
    out["per_tile_by_sm_min_max"] = [float(sm_mean.min()), float(sm_mean.max())]
    order_t = np.argsort(b[full, 7])
    nq = len(order_t) // 4
    out["per_tile_by_launch_quartile"] = [float(pt[order_t[i * nq:(i + 1) * nq]].mean()) for i in range(4)]
    tot = out["prologue"] + out["tile_loop"] + out["drain"] + out["reduction_and_exit"] + out["gap_to_next_cta"]
    out["fixed_share"] = float(1.0 - out["tile_loop"] / tot)
    res[name] = out
    print(name, json.dumps(out))


for name in ("forward", "reverse"):
    buf = torch.zeros(8 * max(n_items, 1) + 64, dtype=torch.int32, device="cuda")
    set_trace(C.c_void_p(buf.data_ptr()))
    for _ in range(3):
        if name == "forward":
            ev.forward(tr.params)
        else:
            ev.backward(ubar, tr.params, g, accumulate=False)
    torch.cuda.synchronize()
    set_trace(None)
    analyse(buf, name)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "item_cost_trace" + ("_small" if kw else "") + ".json"), "w"), indent=1)
