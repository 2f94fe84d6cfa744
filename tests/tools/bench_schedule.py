"""BASELINE config 3 (Burgers, 15x15 subdomains, 200x200 points, FCN [2,16,1]) across a FULL LineScheduler schedule:
training steps/s including every active-set change; the reference recompiles its XLA step at
each change (fbpinns/trainers.py:646-653), here the update inputs are rebuilt on the device and the CUDA graph is
re-captured.  Prints one JSON line (informational, not the headline).

    python tests/tools/bench_schedule.py [--steps 3000] [--no-graph]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbpinns_b200 import configs                            # noqa: E402
from fbpinns_b200.trainers import FBPINNTrainer             # noqa: E402
from fbpinns_b200.util.logger import logger                 # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    logger.setLevel("ERROR")
    c = configs.cfg3_burgers(n_steps=args.steps, use_cuda_graph=not args.no_graph, summary_freq=10 ** 9, test_freq=10 ** 9)
    tr = FBPINNTrainer(c)
    tr.setup()
    scheduler = c.scheduler(all_params=tr.all_params, n_steps=c.n_steps, **c.scheduler_kwargs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t_rebuild, n_rebuild, loss = 0.0, 0, None
    for i, active in enumerate(scheduler):
        if active is not None:
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            tr.set_active(active, i)
            torch.cuda.synchronize()
            t_rebuild += time.perf_counter() - t1
            n_rebuild += 1
        loss = tr.step()
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print(json.dumps({"metric": "train_steps_per_sec_full_schedule", "config": "cfg3 Burgers 15x15, 200x200, LineScheduler",
                      "steps": args.steps, "value": args.steps / total, "unit": "steps/s", "total_s": total,
                      "active_set_changes": n_rebuild, "rebuild_total_s": t_rebuild,
                      "rebuild_ms_each": 1e3 * t_rebuild / max(n_rebuild, 1),
                      "steps_per_s_excluding_rebuilds": args.steps / max(total - t_rebuild, 1e-9),
                      "loss_last": float(loss.item()), "cuda_graph": not args.no_graph}))


if __name__ == "__main__":
    main()
