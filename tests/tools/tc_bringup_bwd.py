"""Hardware bring-up of the tcgen05 reverse kernel (kernel="tensor-full"): gradients against the tiled (FFMA2) reverse
kernel on a reduced cfg 5 with partial tiles and at the full cfg 5 size, kernel timings (CUDA events, L2 flushed) and a
timing breakdown through FBP_TC_DEBUG (2 no MMA, 16 gradient warps idle, 32 no butterfly reduction; those runs produce
wrong results by construction).  Writes gpurun_out/tc_bringup_bwd.json.

    python tests/tools/tc_bringup_bwd.py [--small-only] [--no-breakdown]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fbpinns_b200 import configs, _lib                                   # noqa: E402
from fbpinns_b200._lib import ptr, stream_ptr, check                     # noqa: E402
from fbpinns_b200.engine import Plan, ConstraintEvaluator, unpack_params  # noqa: E402
from fbpinns_b200.trainers import FBPINNTrainer                          # noqa: E402

res = {}


def rel(a, b):
    return float((a - b).abs().max().item() / b.abs().max().clamp_min(1e-30).item())


def run(tag, kw, reps, breakdown):
    lib = _lib.load()
    os.environ["FBP_TC_DEBUG"] = "0"
    c = configs.cfg5_poisson(device="cuda:0", kernel="tiled", use_cuda_graph=False, **kw)
    tr = FBPINNTrainer(c)
    tr.setup()
    tr.set_active(np.ones(tr.all_params["static"]["decomposition"]["m"], dtype=int))
    ev = tr.inputs.evaluators[0]
    ev = getattr(ev, "ev", ev)
    ev.set_affine(None)          # compare the subdomain kernels only (the tensor evaluator below carries no affine operator either)
    plan_t = Plan(ev.plan.layer_sizes, ev.plan.jet, kernel="tensor-full")
    assert plan_t.kernel == "tensor-full", "no tensor instance for this plan"
    ev_t = ConstraintEvaluator(plan_t, ev.takes, ev.x, tr.dd)
    assert ev_t.cache is None
    torch.manual_seed(0)
    ubar = torch.randn(ev.takes.n, ev.V, device="cuda")
    m_act = ev.takes.m_active
    g_ref = torch.zeros((m_act, tr.params.shape[1]), device="cuda")
    g_t = torch.full_like(g_ref, float("nan"))
    ev.forward(tr.params)
    ev.backward(ubar, tr.params, g_ref, accumulate=False)
    ev_t.forward(tr.params)
    ev_t.backward(ubar, tr.params, g_t, accumulate=False)
    torch.cuda.synchronize()
    res[f"{tag}_pairs"] = int(ev.takes.s)
    res[f"{tag}_grad_nan"] = int(torch.isnan(g_t).sum().item())
    res[f"{tag}_grad_rel"] = rel(g_t, g_ref)
    names = ["W0", "b0", "W1", "b1", "W2", "b2"]
    for (gw, gb), (rw, rb), i in zip(unpack_params(plan_t, g_t), unpack_params(plan_t, g_ref), range(3)):
        res[f"{tag}_{names[2 * i]}_rel"] = rel(gw, rw)
        res[f"{tag}_{names[2 * i + 1]}_rel"] = rel(gb, rb)
    print(json.dumps(res), flush=True)

    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")

    def k_bwd(e, g):
        tv = e.takes.view()
        check(lib.fbp_backward(e.plan.handle, C.byref(tv), ptr(e.x), ptr(tr.params), ptr(tr.dd.sub_static), ptr(e.grow),
                               ptr(g), 0, ptr(e.gpart), ptr(e.scratch), e.scratch_floats, ptr(e.cache), stream_ptr()),
              "fbp_backward")

    def timed(e, g):
        k_bwd(e, g)
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); k_bwd(e, g); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        return round(float(np.mean(ts)), 4)

    res[f"{tag}_tiled_bwd_ms"] = timed(ev, g_ref)
    res[f"{tag}_tensor_bwd_ms"] = timed(ev_t, g_t)
    if breakdown:
        for dbg in ((1, 4, 8, 32, 33) if os.environ.get("FBP_TC_BWD", "1") == "2" else (16, 2, 18, 32, 50)):
            os.environ["FBP_TC_DEBUG"] = str(dbg)
            res[f"{tag}_tensor_bwd_dbg{dbg}_ms"] = timed(ev_t, g_t)
        os.environ["FBP_TC_DEBUG"] = "0"
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    try:
        run("small", dict(configs.SMALL["cfg5"], n_sub=(5, 4), n_pts=(160, 136)), 2, False)
        if "--small-only" not in sys.argv:
            run("full", {}, 5, "--no-breakdown" not in sys.argv)
    finally:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        json.dump(res, open(os.path.join(d, "tc_bringup_bwd.json"), "w"), indent=1)
