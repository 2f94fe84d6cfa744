"""GPU parity of the hot path against the oracle (float64 truth) on reduced variants of every BASELINE config:
ujs (unconstrained and through the constraining operator), loss, parameter gradients, within the north-star
tolerance of 1e-5 relative to each quantity's max magnitude (float32 arithmetic on the device)."""
import numpy as np
import pytest
import torch

from fbpinns_b200 import configs
from fbpinns_b200.engine import subdomain_sum, unpack_params, Plan
from fbpinns_b200.jets import JetSpec
from fbpinns_b200.trainers import UpdateStep
from fbpinns_b200.engine import PackedAdam
import common

pytestmark = pytest.mark.gpu

TOL = 1e-5          # north star: u/ujs, loss and gradients within 1e-5 relative (fp32)
NAMES = ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]


def _kernels_for(name):
    return ["generic", "auto"]


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("kernel", ["generic", "auto"])
def test_ujs_match_oracle(name, kernel):
    import gpu_common
    k = common.make_case(configs.CONFIGS[name](**configs.SMALL[name]), seed=0)
    dd, inp, params = gpu_common.device_case(k, kernel=kernel)
    dev = params.device
    all_params = _device_all_params(k, dev)
    for ic, ev in enumerate(inp.evaluators):
        jet = ev.plan.jet
        if ev.takes.n == 0:
            continue
        ujets = ev.forward(params)
        torch.cuda.synchronize()
        ref_plain = common.oracle_ujs(k, ic, torch.float64, constrained=False)
        for (iu, p), got, ref in zip(jet.required_ujs, gpu_common.ujets_columns(jet, ujets), ref_plain):
            e = common.rel_err(got, ref[:, 0])
            assert e < TOL, f"{name} constraint {ic} unconstrained d{p}: rel err {e:.2e}"
        # through the constraining operator (host side, torch.func.jvp on the local Taylor model)
        ref_con = common.oracle_ujs(k, ic, torch.float64, constrained=True)
        ujs = jet.ujs_constrained(ujets, inp.constraints[ic][0], k.c.problem.constraining_fn, all_params)
        for (iu, p), got, ref in zip(jet.required_ujs, ujs, ref_con):
            e = common.rel_err(got.cpu().numpy(), ref)
            assert e < TOL, f"{name} constraint {ic} constrained d{p}: rel err {e:.2e}"


def _device_all_params(k, dev, prob_flat=None):
    st = {}
    for tag, d in k.all_params["static"].items():
        if tag == "decomposition":
            continue
        st[tag] = {kk: (v.to(dev) if torch.is_tensor(v) else v) for kk, v in d.items()}
    tr = {"problem": {kk: torch.as_tensor(v, device=dev) for kk, v in k.prob_trainable.items()}}
    return {"static": st, "trainable": tr}


def _make_step(k, inp, params, dev, graph=False):
    all_params = _device_all_params(k, dev)
    keys = list(k.prob_trainable.keys())
    prob_flat = None
    if keys:
        prob_flat = torch.cat([torch.as_tensor(k.prob_trainable[kk], device=dev).reshape(-1) for kk in keys]).clone()
    adam = PackedAdam(k.m, params.shape[1], 0 if prob_flat is None else prob_flat.numel(), dev, learning_rate=1e-3)
    step = UpdateStep(inp, params, adam, all_params, prob_flat, k.c.problem, use_cuda_graph=graph)
    return step, adam, prob_flat


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("kernel", ["generic", "auto"])
def test_loss_and_grads_match_oracle(name, kernel):
    import gpu_common
    k = common.make_case(configs.CONFIGS[name](**configs.SMALL[name]), seed=0)
    dd, inp, params = gpu_common.device_case(k, kernel=kernel)
    dev = params.device
    step, adam, prob_flat = _make_step(k, inp, params, dev)
    step.grads.zero_()
    loss = step.forward_loss()
    loss.backward()
    torch.cuda.synchronize()
    ref_loss, g_layers, g_prob = common.oracle_loss_and_grads(k, torch.float64)
    assert abs(loss.item() - ref_loss) <= TOL * abs(ref_loss), f"loss {loss.item()} vs {ref_loss}"
    plan = inp.evaluators[0].plan
    got = unpack_params(plan, step.grads[:len(inp.active_ims)].contiguous())
    for l, ((gw, gb), (rw, rb)) in enumerate(zip(got, g_layers)):
        ew, eb = common.rel_err(gw.cpu().numpy(), rw), common.rel_err(gb.cpu().numpy(), rb)
        assert ew < TOL and eb < TOL, f"{name} layer {l}: grad rel err W {ew:.2e} b {eb:.2e}"
    if prob_flat is not None:
        for i, kk in enumerate(k.prob_trainable):
            e = common.rel_err(step.problem_grad().cpu().numpy()[i], g_prob[kk])
            assert e < TOL, f"problem param {kk}: {e:.2e}"


def test_fixed_subdomains_get_no_gradient_and_values_match():
    "active set with fixed (2) subdomains: forward uses them, reverse pass skips them (fbpinns/trainers.py:249-256, 292)"
    import gpu_common
    from fbpinns_b200.schedulers import LineSchedulerRectangularND
    c = configs.cfg3_burgers(n_sub=(5, 5), n_pts=(40, 40), n_steps=40)
    k0 = common.make_case(c, seed=0)
    states = [a.copy() for a in LineSchedulerRectangularND(k0.all_params, 40, point=[0.], iaxis=0) if a is not None]
    active = [a for a in states if (a == 2).any()][0]
    k = common.make_case(c, seed=0, active=active)
    for kernel in ["generic", "auto"]:
        dd, inp, params = gpu_common.device_case(k, kernel=kernel)
        assert len(inp.fixed_ims) > 0
        step, adam, _ = _make_step(k, inp, params, params.device)
        step.grads.zero_()
        loss = step.forward_loss()
        loss.backward()
        ref_loss, g_layers, _ = common.oracle_loss_and_grads(k, torch.float64)
        assert abs(loss.item() - ref_loss) <= TOL * abs(ref_loss)
        got = unpack_params(inp.evaluators[0].plan, step.grads[:len(inp.active_ims)].contiguous())
        for (gw, gb), (rw, rb) in zip(got, g_layers):
            assert common.rel_err(gw.cpu().numpy(), rw) < TOL and common.rel_err(gb.cpu().numpy(), rb) < TOL


def test_mixed_derivative_generic_kernel():
    "u_xy is outside the tiled family: the plan must fall back to the generic kernels and still match the oracle"
    import gpu_common
    from oracle import ref_model
    c = configs.cfg5_poisson(n_sub=(4, 4), n_pts=(40, 40), layer_sizes=(2, 16, 1))
    k = common.make_case(c, seed=3)
    req = ((0, (0, 1)), (0, (1, 1)), (0, ()))
    k.jets = [JetSpec(req, 2, 1)]
    k.jmapss = [ref_model.get_jmaps(req)]
    dd, inp, params = gpu_common.device_case(k, kernel="auto")
    ev = inp.evaluators[0]
    assert not ev.plan.is_fast
    ujets = ev.forward(params)
    ref = common.oracle_ujs(k, 0, torch.float64, constrained=False)
    for (iu, p), got, r in zip(req, gpu_common.ujets_columns(ev.plan.jet, ujets), ref):
        assert common.rel_err(got, r[:, 0]) < TOL, p


def test_partition_of_unity_property_full_size():
    """Size-independent property at BASELINE config 5's full size (64x64 subdomains, 1024x1024 points, where the
    oracle is too slow): with every network constant (zero weights, last bias b) the windowed, normalised sum must
    return u == b at every point and zero first/second derivatives (sum_i w_i / sum_i w_i == 1).  Also checks the
    tiled kernels against the generic ones (two independent CUDA implementations) on random parameters."""
    import gpu_common
    c = configs.cfg5_poisson()
    k = common.Case()
    k.c = c
    sd, _ = c.decomposition.init_params(**c.decomposition_init_kwargs)
    sdom, _ = c.domain.init_params(**c.domain_init_kwargs)
    sp, _ = c.problem.init_params(**c.problem_init_kwargs)
    k.all_params = {"static": {"domain": sdom, "problem": sp, "decomposition": sd}, "trainable": {}}
    k.m, k.xd, k.ud = sd["m"], 2, 1
    k.layer_sizes = list(c.network_init_kwargs["layer_sizes"])
    cons = c.problem.sample_constraints(k.all_params, c.domain, None, "grid", c.ns)
    k.constraints_global = [[t.numpy() for t in cons[0][:-1]]]
    k.x_batch_global = k.constraints_global[0][0]
    k.offsets = np.array([0])
    k.jets = [JetSpec(cons[0][-1], 2, 1)]
    k.active = np.ones(k.m, dtype=int)
    assert len(k.constraints_global[0]) == 2           # points + the static source term
    rng = np.random.default_rng(0)
    from oracle import ref_model
    k.layers = ref_model.init_fcn_params(rng, k.m, k.layer_sizes)
    dd, inp, params = gpu_common.device_case(k, kernel="auto")
    ev = inp.evaluators[0]
    t = ev.takes
    assert t.n == 1024 * 1024 and t.m_all == 4096 and t.s == 8726116
    # constant networks
    P = params.shape[1]
    pc = torch.zeros_like(params)
    pc[:, P - 1] = 0.75
    u = ev.forward(pc)
    torch.cuda.synchronize()
    assert torch.allclose(u[:, 0], torch.full_like(u[:, 0], 0.75), atol=2e-6)
    scale = (np.pi / (2.9 / 63 / 2)) ** 2          # magnitude of the window's second derivative
    assert float(u[:, 1:].abs().max()) < 1e-5 * scale
    # tiled vs generic on random parameters
    u_fast = ev.forward(params).clone()
    if ev.plan.is_fast:
        from fbpinns_b200.engine import ConstraintEvaluator
        gplan = Plan(k.layer_sizes, k.jets[0], kernel="generic")
        gev = ConstraintEvaluator(gplan, t, ev.x, dd)
        u_gen = gev.forward(params)
        for c_ in range(u_gen.shape[1]):
            e = common.rel_err(u_fast[:, c_].cpu().numpy(), u_gen[:, c_].cpu().numpy())
            assert e < TOL, (c_, e)


@pytest.mark.parametrize("name", ["cfg2", "cfg4", "cfg5"])
def test_activation_cache_matches_recompute(name):
    """tiled reverse kernel: TMA-loaded activation cache (forward saves the hidden jets) vs recompute path — same
    gradients up to float32 summation order (both are checked against the oracle elsewhere)."""
    import gpu_common
    from fbpinns_b200.engine import ConstraintEvaluator
    small = dict(configs.SMALL[name])
    if name == "cfg5":
        small.update(n_sub=(5, 4), n_pts=(160, 136))      # full 128-point tiles + partial tails in every subdomain
    k = common.make_case(configs.CONFIGS[name](**small), seed=4)
    dd, inp, params = gpu_common.device_case(k, kernel="auto")
    ev = inp.evaluators[0]
    if ev.plan.is_fast and ev.plan.cache_per_pair == 0:
        pytest.skip("this plan's reverse kernel recomputes the hidden layer (no activation cache)")
    assert ev.plan.is_fast and ev.cache is not None and ev.plan.cache_per_pair > 0
    ev2 = ConstraintEvaluator(ev.plan, ev.takes, ev.x, dd, activation_cache=False)
    assert ev2.cache is None
    torch.manual_seed(0)
    ubar = torch.randn(ev.takes.n, ev.V, device=params.device)
    m_act = max(len(inp.active_ims), 1)
    g1 = torch.zeros((m_act, params.shape[1]), device=params.device)
    g2 = torch.zeros_like(g1)
    u1 = ev.forward(params)
    ev.backward(ubar, params, g1, accumulate=False)
    u2 = ev2.forward(params)
    ev2.backward(ubar, params, g2, accumulate=False)
    torch.cuda.synchronize()
    assert torch.equal(u1, u2)
    assert torch.isfinite(g1).all()
    # 5e-6: where the forward kernel is the tensor-core one, the cached hidden jets are 3xTF32 results (measured 7e-7
    # from the FP32 ones) while the recompute path is pure FP32; each path is held to 1e-5 against the oracle elsewhere
    assert common.rel_err(g1.cpu().numpy(), g2.cpu().numpy()) < 5e-6
    # a second step with different parameters must refresh the cache (no stale activations)
    p2 = params * 1.01
    ev.forward(p2); ev.backward(ubar, p2, g1, accumulate=False)
    ev2.forward(p2); ev2.backward(ubar, p2, g2, accumulate=False)
    assert common.rel_err(g1.cpu().numpy(), g2.cpu().numpy()) < 5e-6


def test_multilevel_decomposition_npou2_matches_oracle():
    """MultilevelRectangularDecompositionND (two partitions of unity, fbpinns/decompositions.py:338-375): per-level
    N/D quotient, sum over levels, /npou (fbpinns/trainers.py:163-170) — ujs, loss and gradients vs the oracle."""
    import gpu_common
    from fbpinns_b200 import decompositions
    from fbpinns_b200.constants import get_subdomain_ws
    xs1 = [np.linspace(-1, 1, 3), np.linspace(0, 1, 2)]
    xs2 = [np.linspace(-1, 1, 5), np.linspace(0, 1, 4)]
    c = configs.cfg3_burgers(n_sub=(3, 2), n_pts=(40, 30), line_scheduler=False)
    c.decomposition = decompositions.MultilevelRectangularDecompositionND
    c.decomposition_init_kwargs = dict(subdomain_xss=[xs1, xs2],
                                       subdomain_wss=[get_subdomain_ws(xs1, 2.9), get_subdomain_ws(xs2, 2.9)], unnorm=(0., 3.))
    k = common.make_case(c, seed=7, multilevel=True)
    assert k.ui["takess"][0][4] == 2
    for kernel in ["generic", "auto"]:
        dd, inp, params = gpu_common.device_case(k, kernel=kernel)
        ev = inp.evaluators[0]
        assert ev.takes.npou == 2 and ev.takes.q > ev.takes.n
        ujets = ev.forward(params)
        ref = common.oracle_ujs(k, 0, torch.float64, constrained=False)
        for (iu, p), got, r in zip(ev.plan.jet.required_ujs, gpu_common.ujets_columns(ev.plan.jet, ujets), ref):
            assert common.rel_err(got, r[:, 0]) < TOL, (kernel, p)
        step, adam, _ = _make_step(k, inp, params, params.device)
        assert step.affine[0] is None            # npou > 1: constraining goes through the generic path
        step.grads.zero_()
        loss = step.forward_loss()
        loss.backward()
        ref_loss, g_layers, _ = common.oracle_loss_and_grads(k, torch.float64)
        assert abs(loss.item() - ref_loss) <= TOL * abs(ref_loss)
        got = unpack_params(ev.plan, step.grads[:len(inp.active_ims)].contiguous())
        for (gw, gb), (rw, rb) in zip(got, g_layers):
            assert common.rel_err(gw.cpu().numpy(), rw) < TOL and common.rel_err(gb.cpu().numpy(), rb) < TOL


@pytest.mark.parametrize("name,small", [
    ("cfg5", dict(n_sub=(4, 4), n_pts=(48, 48), layer_sizes=(2, 64, 64, 1))),      # sweep network of config 5
    ("cfg4", dict(n_sub=(3, 3, 3), n_pts=(12, 12, 12), layer_sizes=(3, 64, 64, 1))),  # config 4's true layer sizes
    ("cfg3", dict(n_sub=(4, 4), n_pts=(40, 40), layer_sizes=(2, 64, 1), line_scheduler=False)),
])
def test_width64_kernel_instances_match_oracle(name, small):
    "H = 64 instances of the tiled family (32-point reverse tiles, 4 weight-gradient quadrants): loss and gradients"
    import gpu_common
    k = common.make_case(configs.CONFIGS[name](**small), seed=11)
    dd, inp, params = gpu_common.device_case(k, kernel="auto")
    assert all(ev.plan.is_fast for ev in inp.evaluators)
    step, adam, _ = _make_step(k, inp, params, params.device)
    step.grads.zero_()
    loss = step.forward_loss()
    loss.backward()
    torch.cuda.synchronize()
    ref_loss, g_layers, _ = common.oracle_loss_and_grads(k, torch.float64)
    assert abs(loss.item() - ref_loss) <= TOL * abs(ref_loss), (loss.item(), ref_loss)
    got = unpack_params(inp.evaluators[0].plan, step.grads[:len(inp.active_ims)].contiguous())
    for l, ((gw, gb), (rw, rb)) in enumerate(zip(got, g_layers)):
        ew, eb = common.rel_err(gw.cpu().numpy(), rw), common.rel_err(gb.cpu().numpy(), rb)
        assert ew < TOL and eb < TOL, f"{name} layer {l}: grad rel err W {ew:.2e} b {eb:.2e}"


def _full_size_cfg5(kernel="auto"):
    "the trainer's own setup of BASELINE config 5 at full size (device takes, evaluator, packed parameters)"
    from fbpinns_b200.trainers import FBPINNTrainer
    from fbpinns_b200.util.logger import logger
    logger.setLevel("WARNING")
    tr = FBPINNTrainer(configs.cfg5_poisson(device="cuda:0", kernel=kernel, use_cuda_graph=False)).setup()
    tr.set_active(np.ones(tr.dd.m, dtype=int))
    ev = tr.inputs.evaluators[0]
    ev.set_affine(None)                 # the subdomain kernels and the quotient rule only (constraining is tested elsewhere)
    return tr, ev


def test_full_size_cfg5_sampled_points_match_oracle():
    """BASELINE config 5 at FULL size (4096 subdomains, 1024 x 1024 points, 8 726 116 pairs — where the 4096-item launch
    order, the int32 offsets and the partial tiles live): the unconstrained ujs of ~500 sampled points, each with ALL its
    pairs, against the float64 oracle (reference-order takes restricted to those points)."""
    from oracle import ref_model
    from fbpinns_b200.engine import unpack_params
    tr, ev = _full_size_cfg5()
    t, jet = ev.takes, ev.plan.jet
    assert t.s == 8726116 and t.n == 1024 * 1024
    ujets = ev.forward(tr.params)
    torch.cuda.synchronize()
    rng = np.random.default_rng(0)
    n = t.n
    pts = np.unique(np.concatenate([rng.integers(0, n, 480), [0, 1023, n - 1024, n - 1],       # corners
                                    1024 * rng.integers(0, 1024, 8), rng.integers(0, 1024, 8)]))  # edges
    m_take, n_take, p_take, np_take, npou = t.reference_arrays()
    sel = np.isin(n_take, pts)
    inv = -np.ones(n, dtype=np.int64)
    inv[pts] = np.arange(len(pts))
    rows = np.unique(p_take[sel])                               # npou == 1: one row per point
    rinv = -np.ones(len(np_take), dtype=np.int64)
    rinv[rows] = np.arange(len(rows))
    takes_s = (m_take[sel], inv[n_take[sel]], rinv[p_take[sel]], inv[np_take[rows]], npou)
    dt = torch.float64
    layers = [(w.cpu().to(dt), b.cpu().to(dt)) for w, b in unpack_params(ev.plan, tr.params)]
    dparams = [torch.as_tensor(np.asarray(p.cpu() if torch.is_tensor(p) else p), dtype=dt)
               for p in tr.all_params["static"]["decomposition"]["subdomain"]["params"]]
    ims = torch.as_tensor(t.sub_ids.cpu().numpy().astype(np.int64))
    dc = {"subdomain": {"params": [p[ims] for p in dparams]}}
    lc = [(w[ims], b[ims]) for w, b in layers]
    x = ev.x.cpu().to(dt)[torch.as_tensor(pts)]
    ref = ref_model.fbpinn_forward(dc, lc, x, takes_s, ref_model.get_jmaps(jet.required_ujs), None, None)
    got = jet.ujs_plain(ujets)
    for (iu, p), g_, r_ in zip(jet.required_ujs, got, ref):
        # the north-star metric over the WHOLE field: error of the sampled points relative to the field's largest value
        scale = float(g_.abs().max())
        err = float(np.max(np.abs(g_[torch.as_tensor(pts, device=g_.device)].cpu().numpy() - r_.numpy()))) / scale
        assert err < TOL, (p, err)


def test_full_size_cfg5_sampled_subdomain_gradients_match_oracle():
    """Full-size config 5: the gradients of 8 sampled subdomains (corner, edge, interior, and the ones with the largest
    gradient) for a random cotangent of the jets, against the float64 oracle evaluated on all pairs of those subdomains."""
    from oracle import ref_model
    from fbpinns_b200.engine import unpack_params
    tr, ev = _full_size_cfg5()
    t, jet = ev.takes, ev.plan.jet
    torch.manual_seed(0)
    ubar = torch.randn(t.n, ev.V, device="cuda")
    g = torch.full((t.m_active, tr.params.shape[1]), float("nan"), device="cuda")
    ev.forward(tr.params)
    ev.backward(ubar, tr.params, g, accumulate=False)
    torch.cuda.synchronize()
    assert torch.isfinite(g).all()
    got = [(w.cpu().double().numpy(), b.cpu().double().numpy()) for w, b in unpack_params(ev.plan, g)]
    grow = ev.grow.cpu().double()
    sub_off = t.sub_off.cpu().numpy()
    sp_point, sp_row = t.spair_point.cpu().numpy(), t.spair_row.cpu().numpy()
    sub_ids = t.sub_ids.cpu().numpy()
    x = ev.x.cpu().double()
    dt = torch.float64
    layers = [(w.cpu().to(dt), b.cpu().to(dt)) for w, b in unpack_params(ev.plan, tr.params)]
    dparams = [torch.as_tensor(np.asarray(p.cpu() if torch.is_tensor(p) else p), dtype=dt)
               for p in tr.all_params["static"]["decomposition"]["subdomain"]["params"]]
    jmaps = ref_model.get_jmaps(tuple((0, p) for p in jet.comps))
    gmax = np.max(np.abs(got[1][0]).reshape(t.m_active, -1), axis=1)
    pick = list(dict.fromkeys([int(v) for v in np.argsort(-gmax)[:3]] + [0, 63, t.m_active - 1, 2093, 1105]))
    for sp in pick:
        im = int(sub_ids[sp])
        a, b = int(sub_off[sp]), int(sub_off[sp + 1])
        xs = x[sp_point[a:b]]
        rows = torch.as_tensor(sp_row[a:b], dtype=torch.long)
        leaves = [(w[im].clone().requires_grad_(True), bb[im].clone().requires_grad_(True)) for w, bb in layers]
        s = b - a
        ps_take = [p[im].expand(s, *p.shape[1:]) for p in dparams]
        lay_take = [(w.expand(s, *w.shape), bb.expand(s, *bb.shape)) for w, bb in leaves]
        jets = torch.cat(ref_model.get_ujs(xs, jmaps, lambda xb: (ref_model.model_inner(ps_take, lay_take, xb)[0], ())), dim=1)
        gr = torch.autograd.grad((grow[rows] * jets).sum(), [t_ for wb in leaves for t_ in wb])
        for l in range(len(layers)):
            for which in (0, 1):
                ref = gr[2 * l + which].numpy()
                # relative to the largest entry of this parameter group over ALL subdomains (the north-star metric)
                scale = float(np.max(np.abs(got[l][which])))
                err = float(np.max(np.abs(got[l][which][sp] - ref))) / scale
                assert err < TOL, (sp, l, which, err)


def test_step_gradients_are_bit_reproducible():
    """No float atomics on the tiled / tensor path: every partial sum is reduced in a fixed order, so two evaluations of the
    same step give bit-identical jets and gradients (fbp_fast.cuh / fbp_tc_bwd2.cuh)."""
    import gpu_common
    k = common.make_case(configs.cfg5_poisson(n_sub=(6, 5), n_pts=(150, 130)), seed=2)
    for kernel in ["auto", "tiled"]:
        dd, inp, params = gpu_common.device_case(k, kernel=kernel)
        ev = inp.evaluators[0]
        torch.manual_seed(1)
        ubar = torch.randn(ev.takes.n, ev.V, device=params.device)
        outs = []
        for rep in range(3):
            g = torch.full((max(len(inp.active_ims), 1), params.shape[1]), float("nan"), device=params.device)
            u = ev.forward(params).clone()
            ev.backward(ubar, params, g, accumulate=False)
            torch.cuda.synchronize()
            outs.append((u, g))
        for u, g in outs[1:]:
            assert torch.equal(u, outs[0][0]) and torch.equal(g, outs[0][1]), kernel


def test_direct_gradient_write_matches_partial_buffer_path():
    """FBP_BWD_DIRECT (one work item per active subdomain: the reverse kernels write the gradient rows themselves, no partial
    buffer, no reduction pass) gives bit-identical gradients, for the tensor and the tiled family; the flag is refused when the
    work list has split subdomains."""
    import gpu_common
    from fbpinns_b200._lib import FbpError
    k = common.make_case(configs.cfg5_poisson(n_sub=(32, 32), n_pts=(256, 256)), seed=4)
    for kernel in ["auto", "tiled"]:
        dd, inp, params = gpu_common.device_case(k, kernel=kernel)
        ev = inp.evaluators[0]
        assert ev.takes.one_item_per_sub and ev.supports_direct_grads()
        torch.manual_seed(3)
        ubar = torch.randn(ev.takes.n, ev.V, device=params.device)
        ev.forward(params)
        g_ref = torch.full((len(inp.active_ims), params.shape[1]), float("nan"), device=params.device)
        ev.backward(ubar, params, g_ref, accumulate=False)
        g_dir = torch.full_like(g_ref, float("nan"))
        ev.direct_grads = True
        try:
            ev.backward(ubar, params, g_dir, accumulate=True)       # the flag overrides: rows are written, not added
        finally:
            ev.direct_grads = False
        torch.cuda.synchronize()
        assert torch.equal(g_dir, g_ref), kernel
    # a small problem: subdomains are split into several work items, the evaluator refuses the flag
    k2 = common.make_case(configs.cfg5_poisson(n_sub=(4, 4), n_pts=(128, 128)), seed=4)
    dd, inp, params = gpu_common.device_case(k2, kernel="auto")
    ev = inp.evaluators[0]
    assert not ev.takes.one_item_per_sub
    ev.forward(params)
    ev.direct_grads = True
    with pytest.raises(FbpError):
        ev.backward(torch.zeros(ev.takes.n, ev.V, device=params.device), params,
                    torch.zeros((len(inp.active_ims), params.shape[1]), device=params.device))
    ev.direct_grads = False
