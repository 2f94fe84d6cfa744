"""Staged multi-GPU smoke run (torchrun). Every stage is logged to gpurun_out/multi_debug_rank<r>.log as it
completes, so that a hang still leaves evidence.  Usage:
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29812 tests/multi_gpu_debug.py [cfg5|cfg2|cfg3]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", f"multi_debug_rank{rank}.log"), "w")
T0 = time.time()


def log(*a):
    print(f"[{time.time() - T0:7.2f}s r{rank}]", *a, file=LOG, flush=True)
    print(f"[{time.time() - T0:7.2f}s r{rank}]", *a, flush=True)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    log("process group up")
    t = torch.ones(4, device="cuda") * rank
    dist.all_reduce(t)
    log("all_reduce ok", t.tolist())
    inp = torch.arange(3, dtype=torch.float32, device="cuda") + 10 * rank
    splits = [1, 2] if rank == 0 else [2, 1]
    if world == 2:
        out = torch.empty(3, device="cuda")
        dist.all_to_all_single(out, inp, output_split_sizes=splits, input_split_sizes=splits)
        log("all_to_all_single ok", out.tolist())
    from fbpinns_b200 import configs
    from fbpinns_b200.trainers import FBPINNTrainer
    from fbpinns_b200.parallel import shard_trainer
    from fbpinns_b200.util.logger import logger
    logger.setLevel("WARNING")

    def make(graph):
        dev = f"cuda:{local}"
        if name == "cfg5":
            return configs.cfg5_poisson(n_sub=(8, 6), n_pts=(96, 72), device=dev, use_cuda_graph=graph)
        if name == "cfg2":
            return configs.cfg2_harmonic_oscillator_inverse(n_sub=10, n_pts=120, device=dev, use_cuda_graph=graph)
        return configs.cfg3_burgers(n_sub=(6, 5), n_pts=(48, 40), line_scheduler=False, device=dev, use_cuda_graph=graph)
    n_steps = 10
    ref = FBPINNTrainer(make(False)).setup()
    ref.set_active(np.ones(ref.dd.m, dtype=int))
    ref_losses = [float(ref.step().item()) for _ in range(n_steps)]
    log("single-GPU reference trajectory", ["%.6e" % v for v in ref_losses[:3]], "...", "%.6e" % ref_losses[-1])
    for graph in (False, True):
        tr = shard_trainer(FBPINNTrainer(make(graph)), rank, world).setup()
        log(f"graph={graph}: setup done")
        tr.set_active(np.ones(tr.dd.m, dtype=int))
        t0 = tr.inputs.takess[0]
        log(f"graph={graph}: set_active done: local points {t0.n}, pairs {t0.s}, subdomains {t0.m_all}, "
            f"send {tr.inputs.evaluators[0].halo.send_counts} recv {tr.inputs.evaluators[0].halo.recv_counts} "
            f"weights {tr.inputs.weights}")
        losses = []
        for i in range(n_steps):
            losses.append(float(tr.step().item()))
            log(f"graph={graph}: step {i} loss {losses[-1]:.6e} (ref {ref_losses[i]:.6e}) captured={tr.update.graph is not None}")
        lo, hi = tr.shard.block(tr.dd.m)
        perr = float((tr.params[lo:hi] - ref.params[lo:hi]).abs().max() / ref.params.abs().max())
        rel = float(np.max(np.abs(np.array(losses) - np.array(ref_losses)) / np.abs(np.array(ref_losses))))
        log(f"graph={graph}: max rel loss dev {rel:.3e}, own-block param dev {perr:.3e}")
    dist.barrier()
    log("done")
    LOG.flush()
    os._exit(0)          # NCCL teardown with captured graphs alive can hang (see bench.py)


if __name__ == "__main__":
    try:
        main()
    except Exception as e:       # make the failure visible in the per-rank log before dying
        import traceback
        log("EXCEPTION", repr(e))
        traceback.print_exc(file=LOG)
        LOG.flush()
        os._exit(1)
