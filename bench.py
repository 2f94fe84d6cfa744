#!/usr/bin/env python
"""
bench.py — FBPINN training-step throughput on B200 (see BASELINE.json / DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA kernels through the public trainer API)
  python bench.py --impl reference [--steps K] [--warmup W]      CPU arm: the restated reference (oracle/, torch CPU,
                                                                 all host threads) on a bounded sample of the workload

A "step" is one FBPINN_update equivalent (fbpinns/trainers.py:285-296): forward jets of every constraint, the
constraining operator, loss_fn, the reverse pass to per-subdomain gradients and the Adam update, steady state,
active set fixed.  Workload at N=1: BASELINE config 5 (Poisson2D, 64x64 = 4096 subdomains, 1024x1024 collocation
grid, FCN [2,32,32,1], jets u_xx, u_yy + chains), synthetic grid points and U(-1/sqrt(fan_in), ..) parameters.
For N>1 (torchrun, one rank per GPU) the same global problem is sharded by subdomain slabs: strong scaling.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FULL = dict(n_sub=(64, 64), n_pts=(1024, 1024), layer_sizes=(2, 32, 32, 1))
WORKLOAD = "cfg5 Poisson2D 64x64 subdomains, 1024x1024 grid, FCN [2,32,32,1], jets (u_x,u_xx,u_y,u_yy)"


def flops_per_pair(layer_sizes, C):
    "algorithmic forward FLOPs per (point, subdomain) pair: 2 MAC_0 + 2 C sum_{l>=1} MAC_l (SURVEY §8d)"
    macs = [a * b for a, b in zip(layer_sizes[:-1], layer_sizes[1:])]
    return 2 * macs[0] + 2 * C * sum(macs[1:])


# ------------------------------------------------------------------------------------------------ clocks

class ClockSampler:
    "samples nvidia-smi clocks / throttle reasons while the timed region runs"
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thr.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm

def cpu_sample_case(n_sub=(4, 4), n_pts=(64, 64)):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fbpinns_b200 import configs
    import common
    c = configs.cfg5_poisson(n_sub=n_sub, n_pts=n_pts, layer_sizes=FULL["layer_sizes"])
    return common.make_case(c, seed=0), common


def time_cpu_steps(steps, warmup, n_sub=(4, 4), n_pts=(64, 64)):
    """Times `steps` FBPINN_update steps of the restated reference (oracle/, float32, per-pair gathered weights,
    nested jvp, index-add segment sums, Adam) on a sample of the workload with the same points per subdomain."""
    import numpy as np
    import torch
    from oracle import ref_adam, ref_model, ref_step
    k, common = cpu_sample_case(n_sub, n_pts)
    torch.set_num_threads(os.cpu_count() or 1)
    dtype = torch.float32
    ui = k.ui
    decomp_cut = ref_model.cut_decomp(ref_model.to_torch(k.decomp_np, dtype), ui["all_ims"])
    al = [(w[ui["active_ims"]].copy(), b[ui["active_ims"]].copy()) for w, b in k.layers]
    fl = [(w[ui["fixed_ims"]], b[ui["fixed_ims"]]) for w, b in k.layers]
    st = ref_adam.adam_init([t for wb in al for t in wb])
    pairs = int(len(ui["takess"][0][0]))
    times, loss = [], None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        loss, al, _, st = ref_step.update(al, fl, {}, st, decomp_cut, ui["takess"], ui["constraints"], k.jmapss,
                                          k.c.problem.loss_fn, k.c.problem.constraining_fn,
                                          common.oracle_all_params(k, dtype), dtype, learning_rate=1e-3)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    return dict(sec_per_sample_step=sec, pairs=pairs, points=int(k.x_batch_global.shape[0]), loss=loss,
                sample=f"cfg5 sub-grid {n_sub[0]}x{n_sub[1]} subdomains, {n_pts[0]}x{n_pts[1]} points "
                       f"({pairs} pairs, same points per subdomain as the full grid), float32, {steps} steps")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    full_pairs = 8726116
    r = time_cpu_steps(max(1, args.steps), max(1, args.warmup))
    pairs_per_s = r["pairs"] / r["sec_per_sample_step"]
    value = pairs_per_s / full_pairs
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": "train_steps_per_sec", "value": value, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "whole-workload steps/s extrapolated per pair from the bounded sample"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": r["sample"] + "; restated reference (torch CPU), not JAX (jax/optax are not installed)",
                         "sample_step_s": r["sec_per_sample_step"], "pair_evals_per_sec": pairs_per_s},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm

def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from fbpinns_b200 import configs, _lib
    from fbpinns_b200.trainers import FBPINNTrainer
    from fbpinns_b200.engine import fma_peak_tflops
    from fbpinns_b200.util.logger import logger
    logger.setLevel("WARNING")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"
    lib = _lib.load()

    layer_sizes = tuple(int(v) for v in args.layers.split(","))
    kw = dict(n_sub=FULL["n_sub"], n_pts=FULL["n_pts"], layer_sizes=layer_sizes)
    if args.small:
        kw.update(n_sub=(16, 16), n_pts=(256, 256))
    if args.shape:          # debug / profiling shapes, e.g. one rank's share of an 8-GPU run on a single GPU: 8,64,128,1024
        a_, b_, c_, d_ = (int(v) for v in args.shape.split(","))
        kw.update(n_sub=(a_, b_), n_pts=(c_, d_))
        args.small = True   # everything that labels / extrapolates the full workload treats it like --small
    if args.config == "cfg5":
        c = configs.cfg5_poisson(device=str(dev), use_cuda_graph=not args.no_graph, kernel=args.kernel, **kw)
    else:       # the other BASELINE configs at full size, all subdomains active (informational lines, not the headline)
        extra = dict(line_scheduler=False) if args.config == "cfg3" else {}
        c = configs.CONFIGS[args.config](device=str(dev), use_cuda_graph=not args.no_graph, kernel=args.kernel, **extra)
        layer_sizes = tuple(c.network_init_kwargs["layer_sizes"])
        kw["n_pts"] = c.ns[0]
    tr = FBPINNTrainer(c)
    if world > 1:
        from fbpinns_b200.parallel import shard_trainer
        shard_trainer(tr, rank, world)
    tr.setup()
    m = tr.dd.m
    tr.set_active(np.ones(m, dtype=int))
    ev = tr.inputs.evaluators[0]
    ev = ev.ev if hasattr(ev, "ev") else ev          # sharded evaluators wrap the local one
    takes = tr.inputs.takess[0]
    C = ev.plan.jet.C
    n_points_global = int(np.prod(kw["n_pts"]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # internal pre-warm: 3 eager steps + graph capture + 1 replay (real training steps, untimed)
    for _ in range(5):
        tr.step()
    for _ in range(max(args.warmup, 3)):
        tr.step()
    loss_first = float(tr.step().item())

    # ---- timed region: K steps, inputs resident in HBM --------------------------------------------------------
    sampler = ClockSampler(local)
    n0 = lib.fbp_launch_count()
    barrier()
    if rank == 0:                   # one poller only: N concurrent nvidia-smi loops stall the driver and the ranks
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss_t = tr.step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    ms = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    loss_last = float(loss_t.item())
    launches_host = int(lib.fbp_launch_count() - n0)
    per_step = tr.update.kernel_launches_per_step
    gpu_launches = per_step * args.steps if (per_step and tr.update.graph is not None) else launches_host
    steps_per_s = 1e3 / ms_per_step

    # ---- end to end: host (pinned) points -> device, step, loss -> host, through the public API -------------------
    x_host = [b.detach().cpu().pin_memory() for b in tr.point_buffers()]
    h2d = int(sum(x.numel() * 4 for x in x_host))
    for _ in range(3):
        tr.step_from_host(x_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tr.step_from_host(x_host)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_steps_per_s = args.steps / float(t_e2e.item())

    # ---- roofline of the dominant kernels: each timed alone with CUDA events, L2 flushed between launches --------
    import ctypes as Cc
    from fbpinns_b200._lib import ptr, stream_ptr, check
    f_fwd = flops_per_pair(layer_sizes, C)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)     # 256 MB > 126 MB L2
    tv = takes.view()
    ubar = torch.randn(takes.n, ev.V, device=dev)
    grads = tr.update.grads

    def time_kernel(fn, reps):
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return float(np.mean(ts)), float(ts[0])

    def k_fwd():
        check(lib.fbp_forward(ev.plan.handle, Cc.byref(tv), ptr(ev.x), ptr(tr.params), ptr(tr.dd.sub_static),
                              ptr(ev.pair_out), ptr(ev.scratch), ev.scratch_floats, ptr(ev.cache), stream_ptr()), "fbp_forward")

    check(lib.fbp_reduce_backward(ev.plan.handle, Cc.byref(tv), ptr(ubar), ptr(ev.dsum), None, ptr(ev.grow), stream_ptr()), "rb")

    def k_bwd():
        check(lib.fbp_backward(ev.plan.handle, Cc.byref(tv), ptr(ev.x), ptr(tr.params), ptr(tr.dd.sub_static), ptr(ev.grow),
                               ptr(grads), 0, ptr(ev.gpart), ptr(ev.scratch), ev.scratch_floats, ptr(ev.cache), stream_ptr()),
              "fbp_backward")

    reps = max(5, min(args.steps, 20))
    k_fwd(); k_bwd()
    fwd_ms, fwd_best = time_kernel(k_fwd, reps)
    bwd_ms, bwd_best = time_kernel(k_bwd, reps)
    fma_scalar = fma_peak_tflops()
    fma_packed = fma_peak_tflops(packed=True)
    fma_peak = max(fma_scalar, fma_packed)          # the roofline denominator is the better of FFMA and FFMA2
    info = torch.cuda.get_device_properties(dev)
    nominal = info.multi_processor_count * 128 * 2 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
    s_local, s_active = takes.s, takes.s_active
    bwd_tf = 2.0 * f_fwd * s_active / (bwd_ms * 1e-3) / 1e12
    fwd_tf = f_fwd * s_local / (fwd_ms * 1e-3) / 1e12
    # names of the kernels fbp_forward / fbp_backward launch for this plan (the dominant one is the reverse kernel)
    bwd_kernel = {"generic": "generic_backward_kernel", "tiled": "fast_backward_kernel",
                  "tensor": "tc_backward_kernel2" if os.environ.get("FBP_TC_BWD", "2") == "2" else "tc_backward_kernel"}[ev.plan.reverse_family]
    fwd_kernel = {"generic": "generic_forward_kernel", "tiled": "fast_forward_kernel",
                  "tensor": "tc_forward_kernel2" if os.environ.get("FBP_TC_FWD", "2") == "2" else "tc_forward_kernel"}[ev.plan.forward_family]
    # DRAM bytes per launch of THAT kernel from the committed ncu capture (profiles/traffic.json, written by
    # profiles/summarise.py from `ncu --set full`: dram__bytes_read.sum + dram__bytes_write.sum); null if never captured
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and args.config == "cfg5" and not args.small:
        try:
            tj = json.load(open(tpath))
            traffic = next((v for k_, v in tj.items() if k_.split("::")[-1].split("(")[0] == bwd_kernel), None)
        except Exception:
            traffic = None
    peaks = {}
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    # algorithmic HBM bytes per step
    P = tr.params.shape[1]
    hbm_bytes = 4 * (takes.n * 2 + 2 * s_local + 2 * takes.n * C + takes.m_all * P * 10)
    step_flops = 0.0
    for e_ in tr.inputs.evaluators:
        e_ = e_.ev if hasattr(e_, "ev") else e_
        step_flops += 3.0 * flops_per_pair(layer_sizes, e_.plan.jet.C) * e_.takes.s
    step_tf = step_flops * world / (ms_per_step * 1e-3) / 1e12

    # ---- active-set change path: full rebuild of the update inputs on the device (inside tests, takes,
    #      work list, window sums, affine jets) — what costs the reference an O(n m) index pass + an XLA recompile
    graph_replayed = tr.update.graph is not None          # read before the rebuild below discards the captured graph
    plan_info = dict(fwd=ev.plan.forward_family, bwd=ev.plan.reverse_family)
    rebuild_ms = None
    if world == 1 and not args.skip_rebuild:
        # as in a training run, nothing but the trainer holds the previous active set's buffers (median of 3 rebuilds)
        del ev, takes, tv, ubar, grads, k_fwd, k_bwd, flush
        ts = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tr.set_active(np.ones(m, dtype=int))
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        rebuild_ms = sorted(ts)[1]

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload -------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        r = time_cpu_steps(steps=2, warmup=1)
        pps = r["pairs"] / r["sec_per_sample_step"]
        cpu = {"value": pps / 8726116 if not args.small else pps / s_local, "unit": "steps/s", "cores": os.cpu_count() or 1,
               "kind": "port", "sample": r["sample"] + "; restated reference (torch CPU), not JAX",
               "pair_evals_per_sec": pps}

    if rank == 0:
        line = {
            "metric": "train_steps_per_sec", "value": steps_per_s, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": (WORKLOAD if args.config == "cfg5" else f"{args.config} (BASELINE config, all subdomains active)")
                       if not args.small else f"DEBUG cfg5 {kw['n_sub'][0]}x{kw['n_sub'][1]} subdomains "
                                                  f"{kw['n_pts'][0]}x{kw['n_pts'][1]} grid (not the headline workload)",
                       "layers": list(layer_sizes), "subdomains": m, "points": n_points_global,
                       "pairs_this_rank": s_local, "parallelism": f"subdomain-slabs x{world}" if world > 1 else "single GPU",
                       "cuda_graph": graph_replayed,
                       "kernel_family": f"forward {plan_info['fwd']} ({fwd_kernel}), reverse {plan_info['bwd']} ({bwd_kernel})"
                                        + ("; tensor = tcgen05 3xTF32, FP32-equivalent accuracy" if "tensor" in plan_info.values() else ""),
                       "l2": "per-step working set (pair jets 175 MB + indices 105 MB) exceeds the 126 MB L2; "
                             "per-kernel timings flush L2 with a 256 MB write between launches"},
            "ujs_point_evals_per_sec": int(tr.x_batch_global.shape[0]) * steps_per_s,
            "pair_evals_per_sec": s_local * world * steps_per_s,
            "step_tflops_algorithmic": step_tf,
            "loss_first": loss_first, "loss_last": loss_last, "active_set_rebuild_ms": rebuild_ms,
            "clocks": clocks,
            "e2e": {"value": e2e_steps_per_s, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "fp32_fma", "kernel": bwd_kernel,
                         "achieved": bwd_tf, "peak": fma_peak,
                         "unit": "TFLOP/s", "frac": bwd_tf / fma_peak if fma_peak else None,
                         "peak_source": "max of the FFMA and FFMA2 (fma.rn.f32x2) micro-benchmarks on this GPU (fbp_fma_peak / "
                                        "fbp_ffma2_peak); MEASURED_PEAKS.json holds only HBM / bf16-tensor peaks",
                         "peak_ffma": fma_scalar, "peak_ffma2": fma_packed, "peak_nominal": nominal,
                         "launch_ms": bwd_ms, "launch_ms_best": bwd_best, "flops_per_launch": 2.0 * f_fwd * s_active,
                         "traffic": traffic,
                         "forward": {"kernel": fwd_kernel,
                                     "achieved": fwd_tf, "frac": fwd_tf / fma_peak if fma_peak else None,
                                     "launch_ms": fwd_ms, "flops_per_launch": float(f_fwd * s_local),
                                     "note": ("hidden GEMM on the tcgen05 tensor cores in 3xTF32 (FP32-equivalent accuracy); the "
                                              "fraction is still algorithmic FP32 flops over the FP32 FMA peak")
                                             if plan_info["fwd"] == "tensor" else None},
                         "step_frac_of_fp32_peak": step_tf / (fma_peak * world) if fma_peak else None,
                         "hbm_algorithmic_gbs": hbm_bytes / (ms_per_step * 1e-3) / 1e9,
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        # CUDA graphs that captured NCCL kernels must go before the communicator; destroy_process_group() can hang
        # on them, so: drop the graphs, synchronise, barrier, and leave without the NCCL teardown.
        tr.update.graph = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layers", default="2,32,32,1", help="FCN layer sizes (sweep: 2,32,1 / 2,64,64,1)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic", "tiled", "tensor", "tensor-full"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-rebuild", action="store_true")
    ap.add_argument("--shape", default=None, help="cfg5 debug shape 'nsub0,nsub1,npts0,npts1' (labels the line as not the headline)")
    ap.add_argument("--small", action="store_true", help="debug-sized problem (not a valid bench number)")
    ap.add_argument("--config", default="cfg5", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE config; cfg5 is the headline workload, the others are informational")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
