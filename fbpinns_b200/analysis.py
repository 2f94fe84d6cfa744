"""
Inference twin of the hot path (API of fbpinns/analysis.py:49-58 `FBPINN_solution` and of the value-only
`FBPINN_model_jit` used by the reference's test step, fbpinns/trainers.py:314-320, 727-737): evaluates the
constrained FBPINN solution at arbitrary points from an `all_params` tree with the reference's leaf layout
(`layers = [(w (m,out,in), b (m,out)), ...]`).  Same kernels as training at jet order 0, no reverse pass.
"""
import numpy as np
import torch

from .engine import Plan, DeviceTakes, ConstraintEvaluator, pack_params
from .jets import JetSpec
from .trainers import active_set_algebra


def FBPINN_model(c, all_params, active, x_batch, device=None):
    """Returns (u, wp, us): the constrained solution (n, ud), the window sums per (point, pou) row (q, 1) and the
    windowed per-pair values u_i * w_i in the reference's (point-sorted) pair order (s, ud) — the first, second and
    third outputs of the reference's FBPINN_model (fbpinns/trainers.py:126-177).  `active` is the scheduler state
    (0/1/2 per subdomain): as in get_inputs, every subdomain that contains a point takes part in the forward."""
    dev = torch.device(device or c.device)
    dd = c.decomposition._device(all_params, dev)
    ud, xd = all_params["static"]["problem"]["dims"]
    from . import networks
    activation, layer_sizes, layers = networks.kernel_layers(c.network, all_params, dev)
    x = x_batch.to(dev, torch.float32).contiguous()
    _, mc = dd.inside_count(x)
    _, a_ims, f_ims, all_ims, pos = active_set_algebra(np.asarray(active), mc.cpu().numpy())
    plan = Plan(layer_sizes, JetSpec(tuple((iu, ()) for iu in range(ud)), xd, ud), kernel=getattr(c, "kernel", "auto"),
                activation=activation)
    takes = DeviceTakes(dd, x, pos, all_ims, len(a_ims), tile_points=plan.tile_points)
    ev = ConstraintEvaluator(plan, takes, x, dd, activation_cache=False)
    params = pack_params(plan, layers)
    with torch.no_grad():
        u = ev.forward(params)
        # the reference hands the FULL tree to constraining_fn (fbpinns/trainers.py:173); tensors / numpy leaves of the
        # static problem and domain entries go to the device, the kernels' private cache entry is dropped
        def dev_tree(d):
            if isinstance(d, dict):
                return {k: dev_tree(v) for k, v in d.items() if k != "_device_cache"}
            if torch.is_tensor(d):
                return d.to(dev)
            if isinstance(d, np.ndarray) and d.dtype.kind == "f":
                return torch.as_tensor(d, dtype=torch.float32, device=dev)
            return d
        ap = {"static": {k: (dev_tree(d) if k in ("problem", "domain") else d) for k, d in all_params["static"].items()},
              "trainable": {k: (dev_tree(v) if k == "problem" else v) for k, v in all_params["trainable"].items()}}
        u = c.problem.constraining_fn(ap, x, u)
        wp = ev.dsum[:takes.q, 0:1].clone()
        us = ev.pair_values_reference_order()
    return u, wp, us


def FBPINN_solution(c, all_params, active, x_batch):
    "Constrained FBPINN solution at x_batch (fbpinns/analysis.py:49-58)"
    return FBPINN_model(c, all_params, active, x_batch)[0]
