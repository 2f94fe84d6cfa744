"""GPU parity of the tcgen05 ("tensor") kernel family (fbpinns_b200/csrc/fbp_tc*.cuh): 3xTF32 hidden-layer GEMMs on the
tensor cores must stay inside the same 1e-5 bar as the FP32 kernels.

Everything runs unconditionally since the round-2 hardware session (profiles/r2a_tc_bringup.md): the MMA self-test, every
forward instance, both forward kernels (FBP_TC_FWD=1 | 2) and the reverse kernel (`kernel="tensor-full"`)."""
import os

import numpy as np
import pytest
import torch

from fbpinns_b200 import configs, _lib
from fbpinns_b200.engine import unpack_params
import common

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _selftest(variant, seed=0):
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    a = torch.randn((128, 32), generator=g).cuda().contiguous()
    w = (torch.rand((32, 32), generator=g) * 2 - 1).cuda().contiguous()
    out = torch.full((128, 32), float("nan"), device="cuda")
    _lib.check(lib.fbp_tc_selftest(_lib.ptr(a), _lib.ptr(w), _lib.ptr(out), variant, _lib.stream_ptr()), "fbp_tc_selftest")
    torch.cuda.synchronize()
    ref = a.double() @ w.double().T
    scale = (a.double().abs() @ w.double().abs().T).max().item()
    return (out.double() - ref).abs().max().item() / scale


def test_mma_selftest_3xtf32_is_fp32_accurate():
    err = _selftest(0)
    assert err < 2e-6, f"3xTF32 MMA self-test: error {err:.2e} relative to |A||W|"


def test_mma_selftest_single_pass_is_tf32_accurate():
    "sanity of the probe itself: one TF32 pass must be right to ~1e-3 and visibly worse than the split"
    err = _selftest(4)
    assert 1e-6 < err < 3e-3, f"single-pass TF32: error {err:.2e}"


@pytest.mark.parametrize("name", ["cfg2", "cfg5"])
def test_tensor_forward_matches_oracle_and_tiled(name):
    import gpu_common
    cases = [dict(configs.SMALL[name])]                   # the size the oracle-level tests of the other families use
    if name == "cfg5":
        cases.append(dict(configs.SMALL[name], n_sub=(5, 4), n_pts=(160, 136)))   # full 128-pair tiles + partial tails
    for i, small in enumerate(cases):
        k = common.make_case(configs.CONFIGS[name](**small), seed=0)
        dd, inp_t, params = gpu_common.device_case(k, kernel="tensor")
        _, inp_f, _ = gpu_common.device_case(k, kernel="tiled")
        ev_t, ev_f = inp_t.evaluators[0], inp_f.evaluators[0]
        assert ev_t.plan.kernel == "tensor" and ev_t.plan.forward_family == "tensor", "the plan has no tensor instance"
        assert ev_f.plan.forward_family == "tiled"
        u_t = ev_t.forward(params)
        u_f = ev_f.forward(params)
        torch.cuda.synchronize()
        jet = ev_t.plan.jet
        if i == 0:
            ref = common.oracle_ujs(k, 0, torch.float64, constrained=False)
            for (iu, p), got, r in zip(jet.required_ujs, gpu_common.ujets_columns(jet, u_t), ref):
                e = common.rel_err(got, r[:, 0])
                assert e < TOL, f"{name} d{p}: tensor forward rel err {e:.2e}"
        assert common.rel_err(u_t.cpu().numpy(), u_f.cpu().numpy()) < 5e-6
        # the activation cache written for the tiled reverse kernel must hold the same hidden jets
        if ev_t.cache is not None and ev_f.cache is not None:
            assert common.rel_err(ev_t.cache.cpu().numpy(), ev_f.cache.cpu().numpy()) < 5e-6


@pytest.mark.parametrize("kernel", ["tensor", "tensor-full"])
@pytest.mark.parametrize("name", ["cfg2", "cfg5"])
def test_tensor_loss_and_grads_match_oracle(name, kernel):
    import gpu_common
    from test_gpu_forward_backward import _make_step
    k = common.make_case(configs.CONFIGS[name](**configs.SMALL[name]), seed=0)
    dd, inp, params = gpu_common.device_case(k, kernel=kernel)
    assert inp.evaluators[0].plan.kernel == kernel
    step, adam, prob_flat = _make_step(k, inp, params, params.device)
    step.grads.zero_()
    loss = step.forward_loss()
    loss.backward()
    torch.cuda.synchronize()
    ref_loss, g_layers, g_prob = common.oracle_loss_and_grads(k, torch.float64)
    assert abs(loss.item() - ref_loss) <= TOL * abs(ref_loss), f"loss {loss.item()} vs {ref_loss}"
    got = unpack_params(inp.evaluators[0].plan, step.grads[:len(inp.active_ims)].contiguous())
    for l, ((gw, gb), (rw, rb)) in enumerate(zip(got, g_layers)):
        ew, eb = common.rel_err(gw.cpu().numpy(), rw), common.rel_err(gb.cpu().numpy(), rb)
        assert ew < TOL and eb < TOL, f"{name} layer {l}: grad rel err W {ew:.2e} b {eb:.2e}"


def test_tensor_reverse_matches_tiled_with_partial_tiles_and_fixed_subdomains():
    """tensor-full vs tiled gradients on a case with full tiles, partial tails and fixed (forward-only) subdomains"""
    import gpu_common
    from fbpinns_b200.schedulers import LineSchedulerRectangularND
    c = configs.cfg5_poisson(n_sub=(5, 4), n_pts=(160, 136), n_steps=40)
    k0 = common.make_case(c, seed=1)
    states = [a.copy() for a in LineSchedulerRectangularND(k0.all_params, 40, point=[0.], iaxis=0) if a is not None]
    for active in [np.ones(k0.m, dtype=int), [a for a in states if (a == 2).any()][0]]:
        k = common.make_case(c, seed=1, active=active)
        grads = {}
        for kernel in ["tiled", "tensor-full"]:
            dd, inp, params = gpu_common.device_case(k, kernel=kernel)
            ev = inp.evaluators[0]
            torch.manual_seed(0)
            ubar = torch.randn(ev.takes.n, ev.V, device=params.device)
            g = torch.full((max(len(inp.active_ims), 1), params.shape[1]), float("nan"), device=params.device)
            ev.forward(params)
            ev.backward(ubar, params, g, accumulate=False)
            torch.cuda.synchronize()
            grads[kernel] = g.cpu().numpy()
        assert np.isfinite(grads["tensor-full"]).all()
        assert common.rel_err(grads["tensor-full"], grads["tiled"]) < 5e-6


def test_tensor_training_curve_matches_tiled():
    "30 Adam steps of the reduced cfg 5 with either family: same loss curve within 1e-4 relative"
    from fbpinns_b200.trainers import FBPINNTrainer
    losses = {}
    for kernel in ["tiled", "tensor", "tensor-full"]:
        c = configs.cfg5_poisson(device="cuda:0", kernel=kernel, use_cuda_graph=True, **configs.SMALL["cfg5"])
        tr = FBPINNTrainer(c)
        tr.setup()
        tr.set_active(np.ones(tr.all_params["static"]["decomposition"]["m"], dtype=int))
        losses[kernel] = [float(tr.step()) for _ in range(30)]
    a = np.array(losses["tiled"])
    for kernel in ["tensor", "tensor-full"]:
        b = np.array(losses[kernel])
        assert np.all(np.abs(a - b) <= 1e-4 * np.abs(a)), (kernel, a[-3:], b[-3:])


def test_both_forward_kernels_agree(monkeypatch):
    "FBP_TC_FWD=1 (first kernel) and the default software-pipelined kernel: same pair outputs and activation cache"
    import gpu_common
    k = common.make_case(configs.cfg5_poisson(n_sub=(5, 4), n_pts=(160, 136)), seed=0)
    outs = {}
    for v in ("1", "2"):
        monkeypatch.setenv("FBP_TC_FWD", v)
        dd, inp, params = gpu_common.device_case(k, kernel="tensor")
        ev = inp.evaluators[0]
        u = ev.forward(params)
        torch.cuda.synchronize()
        outs[v] = (u.cpu().numpy(), ev.cache.cpu().numpy())
    assert common.rel_err(outs["1"][0], outs["2"][0]) < 5e-6
    assert common.rel_err(outs["1"][1], outs["2"][1]) < 5e-6
