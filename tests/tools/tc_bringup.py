"""Hardware bring-up of the tcgen05 forward kernel: parity against the tiled (FFMA2) forward on a reduced cfg 5 with
partial tiles and at the full cfg 5 size, kernel timings (CUDA events, L2 flushed) and a timing breakdown through the
kernel's debug switches (FBP_TC_DEBUG bits: 2 no MMA, 4 no layer 0 / A stores, 8 no epilogue tanh jets; results of
those runs are wrong by construction and are not compared).  Writes gpurun_out/tc_bringup.json.

    python tests/tools/tc_bringup.py [--small-only] [--no-breakdown]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fbpinns_b200 import configs, _lib                      # noqa: E402
from fbpinns_b200._lib import ptr, stream_ptr, check        # noqa: E402
from fbpinns_b200.trainers import FBPINNTrainer             # noqa: E402

res = {}


def rel(a, b):
    return float((a - b).abs().max().item() / b.abs().max().clamp_min(1e-30).item())


def run(tag, kw, reps, breakdown):
    lib = _lib.load()
    c = configs.cfg5_poisson(device="cuda:0", kernel="tiled", use_cuda_graph=False, **kw)
    tr = FBPINNTrainer(c)
    tr.setup()
    tr.set_active(np.ones(tr.all_params["static"]["decomposition"]["m"], dtype=int))
    ev = tr.inputs.evaluators[0]
    ev = getattr(ev, "ev", ev)
    tv = ev.takes.view()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")

    def k_fwd(cache=True):
        check(lib.fbp_forward(ev.plan.handle, C.byref(tv), ptr(ev.x), ptr(tr.params), ptr(tr.dd.sub_static),
                              ptr(ev.pair_out), ptr(ev.scratch), ev.scratch_floats, ptr(ev.cache) if cache else None,
                              stream_ptr()), "fbp_forward")

    def timed(cache=True):
        k_fwd(cache)
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); k_fwd(cache); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        return round(float(np.mean(ts)), 4)

    def outputs(kernel, nwg, fwd=1):
        os.environ["FBP_TC_NWG"], os.environ["FBP_TC_DEBUG"], os.environ["FBP_TC_FWD"] = str(nwg), "0", str(fwd)
        ev.plan.set_kernel(kernel)
        assert ev.plan.kernel == kernel, f"plan fell back to {ev.plan.kernel}"
        ev.pair_out.fill_(float("nan"))
        if ev.cache is not None:
            ev.cache.fill_(float("nan"))
        k_fwd()
        torch.cuda.synchronize()
        return ev.pair_out.clone(), None if ev.cache is None else ev.cache.clone()

    ref = outputs("tiled", 2)
    res[f"{tag}_pairs"] = int(ev.takes.s)
    res[f"{tag}_tiled_ms"] = timed()
    res[f"{tag}_tiled_nocache_ms"] = timed(False)
    for nwg in (2, 4):
        got = outputs("tensor", nwg)
        res[f"{tag}_nwg{nwg}_nan"] = int(torch.isnan(got[0]).sum().item())
        res[f"{tag}_nwg{nwg}_pair_out_rel"] = rel(got[0], ref[0])
        if ref[1] is not None:
            res[f"{tag}_nwg{nwg}_cache_rel"] = rel(got[1], ref[1])
        res[f"{tag}_nwg{nwg}_ms"] = timed()
        res[f"{tag}_nwg{nwg}_nocache_ms"] = timed(False)
        print(json.dumps(res), flush=True)
        if breakdown:
            for dbg in (2, 6, 14, 8):
                os.environ["FBP_TC_DEBUG"] = str(dbg)
                res[f"{tag}_nwg{nwg}_dbg{dbg}_nocache_ms"] = timed(False)
            os.environ["FBP_TC_DEBUG"] = "0"
            print(json.dumps(res), flush=True)
    # software-pipelined variant (FBP_TC_FWD=2)
    got = outputs("tensor", 4, fwd=2)
    res[f"{tag}_v2_nan"] = int(torch.isnan(got[0]).sum().item())
    res[f"{tag}_v2_pair_out_rel"] = rel(got[0], ref[0])
    if ref[1] is not None:
        res[f"{tag}_v2_cache_rel"] = rel(got[1], ref[1])
    res[f"{tag}_v2_ms"] = timed()
    res[f"{tag}_v2_nocache_ms"] = timed(False)
    os.environ["FBP_TC_FWD"] = "1"
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    try:
        run("small", dict(configs.SMALL["cfg5"], n_sub=(5, 4), n_pts=(160, 136)), 2, False)
        if "--small-only" not in sys.argv:
            run("full", {}, 5, "--no-breakdown" not in sys.argv)
    finally:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        json.dump(res, open(os.path.join(d, "tc_bringup.json"), "w"), indent=1)
