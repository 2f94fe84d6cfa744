"""Device-side counterpart of common.make_case: builds the CUDA evaluators for a Case through the C ABI."""
import numpy as np
import torch

from fbpinns_b200.engine import (Plan, DeviceDecomposition, DeviceTakes, ConstraintEvaluator, pack_params)
from fbpinns_b200.trainers import get_update_inputs


def device_case(k, kernel="auto", dev="cuda:0"):
    dev = torch.device(dev)
    d = k.all_params["static"]["decomposition"]
    dd = DeviceDecomposition(d["subdomain"]["params"], d["subdomain"]["pou"], dev)
    cons_g = [[torch.as_tensor(a, dtype=torch.float32, device=dev).contiguous() for a in con] for con in k.constraints_global]
    xg = torch.as_tensor(k.x_batch_global, dtype=torch.float32, device=dev).contiguous()
    inp = get_update_inputs(k.active, k.all_params, dd, xg, cons_g, k.offsets, k.jets, k.layer_sizes, kernel=kernel)
    layers = [(torch.as_tensor(w, device=dev), torch.as_tensor(b, device=dev)) for w, b in k.layers]
    params = pack_params(inp.evaluators[0].plan, layers)
    return dd, inp, params


def ujets_columns(jet, ujets):
    return [ujets[:, jet.column(iu, p)].detach().cpu().numpy() for iu, p in jet.required_ujs]
