"""GPU parity of the reference's other Network plug-ins (AdaptiveFCN, SIREN, AdaptiveSIREN, FourierFCN;
fbpinns/networks.py:70-194) on the generic kernel family's activation variants (csrc/fbp_generic_act.cu) against the
oracle: unconstrained ujs and the gradients of every parameter leaf (weights, biases, activation parameters) for a
random cotangent, 1e-5 relative.

First run on a B200 in round 2 (5 passed, profiles/r2a_tc_bringup.md); unconditional since.  Their maths is checked on the CPU (tests/test_oracle_and_math.py::
test_activation_jet_formulas_match_autograd) and the oracle / host mirrors against the reference's own network_fn
(tests/test_golden_reference.py::test_network_plugins_match_reference_source)."""
import os

import numpy as np
import pytest
import torch

from fbpinns_b200 import configs
from fbpinns_b200 import networks as N
from fbpinns_b200.engine import DeviceDecomposition, pack_params, unpack_params
from fbpinns_b200.trainers import get_update_inputs
from oracle import ref_model
import common

pytestmark = [pytest.mark.gpu]

TOL = 1e-5
NETS = {"fcn": (N.FCN, "fcn", 0), "adaptive_fcn": (N.AdaptiveFCN, "adaptive_fcn", 1), "siren": (N.SIREN, "siren", 0),
        "adaptive_siren": (N.AdaptiveSIREN, "adaptive_siren", 2), "fourier": (N.FourierFCN, "fourier", 0)}


def _case(name, rng):
    c = configs.cfg3_burgers(n_sub=(4, 3), n_pts=(24, 20), line_scheduler=False)
    k = common.make_case(c, seed=0)
    cls, oname, n_extra = NETS[name]
    hidden = [2, 8, 8, 1]
    n_features = 3
    sizes = ([2 * n_features] + hidden[1:]) if name == "fourier" else hidden
    layers = []
    for fi, fo in zip(sizes[:-1], sizes[1:]):
        v = np.sqrt((6.0 if "siren" in name else 1.0) / fi)
        leaf = [rng.uniform(-v, v, (k.m, fo, fi)), rng.uniform(-v, v, (k.m, fo))]
        leaf += [rng.uniform(0.6, 1.4, (k.m, fo)) for _ in range(n_extra)]
        layers.append(tuple(t.astype(np.float32) for t in leaf))
    static = None
    if name == "fourier":
        static = {"omega": (2 * np.pi * (0.1 + 0.3 * rng.standard_normal((k.m, n_features, 2)))).astype(np.float32)}
    return k, cls, oname, layers, static


@pytest.mark.parametrize("name", list(NETS))
def test_network_plugin_ujs_and_gradients_match_oracle(name):
    rng = np.random.default_rng(11)
    k, cls, oname, layers, static = _case(name, rng)
    dev = torch.device("cuda:0")
    # ---- device side through the public pieces: kernel view of the network, plan with its activation, evaluator
    ap = {"static": {}, "trainable": {"network": {"subdomain": {"layers": [tuple(torch.tensor(t) for t in leaf) for leaf in layers]}}}}
    if static is not None:
        ap["static"]["network"] = {"subdomain": {kk: torch.tensor(v) for kk, v in static.items()}}
    activation, ksizes, klayers = N.kernel_layers(cls, ap, dev)
    d = k.all_params["static"]["decomposition"]
    dd = DeviceDecomposition(d["subdomain"]["params"], d["subdomain"]["pou"], dev)
    cons_g = [[torch.as_tensor(a, dtype=torch.float32, device=dev).contiguous() for a in con] for con in k.constraints_global]
    xg = torch.as_tensor(k.x_batch_global, dtype=torch.float32, device=dev).contiguous()
    inp = get_update_inputs(k.active, k.all_params, dd, xg, cons_g, k.offsets, k.jets, ksizes, kernel="auto", activation=activation)
    ev = inp.evaluators[0]
    assert not ev.plan.is_fast or name == "fcn"
    params = pack_params(ev.plan, klayers)
    ujets = ev.forward(params)
    torch.manual_seed(0)
    ubar = torch.randn_like(ujets)
    grads = torch.full((len(inp.active_ims), params.shape[1]), float("nan"), device=dev)
    ev.backward(ubar, params, grads, accumulate=False)
    torch.cuda.synchronize()

    # ---- oracle (float64): ujs and the gradient of sum(ubar * ujets) with respect to every leaf
    ui = k.ui
    decomp_cut = ref_model.cut_decomp(ref_model.to_torch(k.decomp_np, torch.float64), ui["all_ims"])
    lc = [tuple(torch.tensor(t[ui["all_ims"]], dtype=torch.float64, requires_grad=True) for t in leaf) for leaf in layers]
    sc = None if static is None else {kk: torch.tensor(v[ui["all_ims"]], dtype=torch.float64) for kk, v in static.items()}
    x = torch.as_tensor(ui["constraints"][0][0], dtype=torch.float64)
    ujs = ref_model.fbpinn_forward(decomp_cut, lc, x, ui["takess"][0], k.jmapss[0], None, None, oname, sc)
    jet = ev.plan.jet
    L = 0.0
    for (iu, p), ref in zip(jet.required_ujs, ujs):
        col = jet.column(iu, p)
        e = common.rel_err(ujets[:, col].cpu().numpy(), ref[:, 0].detach().numpy())
        assert e < TOL, f"{name} d{p}: rel err {e:.2e}"
        L = L + (torch.tensor(ubar[:, col].cpu().numpy(), dtype=torch.float64) * ref[:, 0]).sum()
    # cotangents of jet components that are not required outputs do not reach the oracle's loss: zero them on the device
    req_cols = {jet.column(iu, p) for iu, p in jet.required_ujs}
    assert req_cols == set(range(ujets.shape[1])), "the Burgers jets are all required outputs"
    ref_grads = torch.autograd.grad(L, [t for leaf in lc for t in leaf], allow_unused=True)
    got = N.from_kernel_layers(cls, unpack_params(ev.plan, grads))
    if name == "fourier":        # the static feature layer must not receive a gradient
        g0 = unpack_params(ev.plan, grads)[0]
        assert float(g0[0].abs().max()) == 0.0 and float(g0[1].abs().max()) == 0.0
    it = iter(ref_grads)
    for l, leaf in enumerate(got):
        for i, g in enumerate(leaf):
            r = next(it)
            if r is None:            # the last layer's activation parameters are unused (as in the reference)
                assert float(g.abs().max()) == 0.0
                continue
            e = common.rel_err(g.cpu().numpy(), r.numpy())
            assert e < TOL, f"{name} layer {l} leaf {i}: grad rel err {e:.2e}"
