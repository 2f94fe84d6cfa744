"""Golden vectors made by running the REFERENCE'S OWN SOURCE under a numpy shim (tests/golden/make_golden_shim.py):
pins the oracle (CPU) and, through the same fixtures, the CUDA path (GPU)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_takes, ref_model
from fbpinns_b200 import problems, decompositions
from fbpinns_b200.jets import JetSpec, get_jmaps
import common

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from cases import NAMES, case_setup  # noqa: E402


def _load(name):
    g = np.load(os.path.join(HERE, "golden", f"refmodel_{name}.npz"), allow_pickle=True)
    cs = case_setup(name)
    assert np.array_equal(g["x"], cs["x"]) and repr(cs["req"]) == str(g["req_repr"])
    nl = len(cs["layer_sizes"]) - 1
    layers = [(g[f"W{l}"], g[f"b{l}"]) for l in range(nl)]
    return g, cs, layers


@pytest.mark.parametrize("name", NAMES)
def test_decomposition_and_jmaps_match_reference(name):
    g, cs, _ = _load(name)
    xd = len(cs["dkw"]["subdomain_xs"])
    ours = decompositions.RectangularDecompositionND._get_level_params(0, xd, **cs["dkw"])
    orac = ref_takes.level_params(0, xd, cs["dkw"]["subdomain_xs"], cs["dkw"]["subdomain_ws"], cs["dkw"]["unnorm"])
    for i in range(7):
        assert np.array_equal(np.asarray(ours[i]), g[f"level_{i}"]), i      # float64, bit for bit
        assert np.array_equal(np.asarray(orac[i]), g[f"level_{i}"]), i
    sd, _ = decompositions.RectangularDecompositionND.init_params(**cs["dkw"])
    od = ref_takes.rectangular_init_params(**cs["dkw"])
    for i in range(6):
        assert np.array_equal(sd["subdomain"]["params"][i].numpy(), g[f"static_{i}"])
        assert np.array_equal(od["subdomain"]["params"][i], g[f"static_{i}"])
    assert np.array_equal(sd["xmins0"], g["xmins0"]) and np.array_equal(sd["xmaxs0"], g["xmaxs0"])
    assert repr(get_jmaps(cs["req"])) == str(g["jmaps_repr"])
    assert repr(ref_model.get_jmaps(cs["req"])) == str(g["jmaps_repr"])


@pytest.mark.parametrize("name", NAMES)
def test_get_inputs_matches_reference(name):
    g, cs, _ = _load(name)
    decomp = ref_takes.rectangular_init_params(**cs["dkw"])
    takes, all_ims, a_ims, f_ims, active = ref_takes.get_inputs(cs["x"], g["active_in"], decomp)
    assert np.array_equal(active, g["active_out"]) and np.array_equal(all_ims, g["all_ims"])
    for got, nm in zip(takes[:4], ["m_take", "n_take", "p_take", "np_take"]):
        assert np.array_equal(got, g[nm]), nm
    assert takes[4] == int(g["npou"])
    # host-side algebra of the product
    from fbpinns_b200.trainers import active_set_algebra
    counts = ref_takes.inside_mask(decomp, cs["x"], np.arange(decomp["m"])).sum(0)
    act2, a2, f2, all2, pos = active_set_algebra(g["active_in"], counts)
    assert np.array_equal(act2, g["active_out"]) and np.array_equal(all2, g["all_ims"])


def _oracle_model(g, cs, layers, dtype=torch.float64, constrained=True):
    decomp = ref_takes.rectangular_init_params(**cs["dkw"])
    all_ims = g["all_ims"]
    dc = ref_model.cut_decomp(ref_model.to_torch(decomp, dtype), all_ims)
    lc = [(torch.as_tensor(w[all_ims], dtype=dtype), torch.as_tensor(b[all_ims], dtype=dtype)) for w, b in layers]
    takes = (g["m_take"], g["n_take"], g["p_take"], g["np_take"], int(g["npou"]))
    prob = getattr(problems, cs["problem"])
    sp, _ = prob.init_params(**cs["pkw"])
    ap = {"static": {"problem": {k: (v.to(dtype) if torch.is_tensor(v) else v) for k, v in sp.items()}}, "trainable": {}}
    x = torch.as_tensor(cs["x"], dtype=dtype)
    cf = prob.constraining_fn if constrained else None
    return dc, lc, takes, prob, ap, x, cf


@pytest.mark.parametrize("name", NAMES)
def test_oracle_model_values_match_reference(name):
    "FBPINN_model of the reference (its own norm/network/unnorm/window/segment-sum/constraining code) vs the oracle"
    g, cs, layers = _load(name)
    dc, lc, takes, prob, ap, x, cf = _oracle_model(g, cs, layers)
    u, wp, us, ws, us_raw = ref_model.fbpinn_model(dc, lc, x, takes, cf, ap)
    for got, nm in [(u, "u"), (wp, "wp"), (us, "us"), (ws, "ws"), (us_raw, "us_raw")]:
        assert np.allclose(got.numpy(), g[nm], rtol=1e-12, atol=1e-13), nm
    u0 = ref_model.fbpinn_model(dc, lc, x, takes, None, ap)[0]
    assert np.allclose(u0.numpy(), g["u_unconstrained"], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_ujs_match_finite_differences_of_reference(name):
    "nested-jvp ujs of the oracle vs central finite differences of the reference's own FBPINN_model (float64)"
    g, cs, layers = _load(name)
    dc, lc, takes, prob, ap, x, cf = _oracle_model(g, cs, layers)
    ujs = ref_model.fbpinn_forward(dc, lc, x, takes, ref_model.get_jmaps(cs["req"]), cf, ap)
    for j, uj in enumerate(ujs):
        e = common.rel_err(uj.numpy(), g[f"uj_fd_{j}"])
        assert e < 2e-6, (name, j, e)
    # the torch restatement of loss_fn against the reference's loss_fn on the same ujs
    cons = [[x] + [torch.as_tensor(g[f"uj_fd_{j}"]) for j in range(len(cs["req"]))]]
    ours = float(prob.loss_fn(ap, cons))
    assert abs(ours - float(g["loss_on_fd_ujs"])) <= 1e-10 * abs(float(g["loss_on_fd_ujs"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("kernel", ["generic", "auto"])
def test_cuda_path_matches_reference_golden(name, kernel):
    "CUDA kernels through the C ABI on the golden inputs: takes bit-exact, u and ujs within 1e-5"
    from fbpinns_b200.engine import DeviceDecomposition, DeviceTakes, ConstraintEvaluator, Plan, pack_params
    from fbpinns_b200.trainers import active_set_algebra
    g, cs, layers = _load(name)
    dev = torch.device("cuda:0")
    sd, _ = decompositions.RectangularDecompositionND.init_params(**cs["dkw"])
    dd = DeviceDecomposition(sd["subdomain"]["params"], sd["subdomain"]["pou"], dev)
    x = torch.as_tensor(cs["x"], device=dev)
    _, mc = dd.inside_count(x)
    act2, a_ims, f_ims, all_ims, pos = active_set_algebra(g["active_in"], mc.cpu().numpy())
    assert np.array_equal(all_ims, g["all_ims"])
    xd, ud = x.shape[1], 1
    jet = JetSpec(cs["req"], xd, ud)
    plan = Plan(cs["layer_sizes"], jet, kernel=kernel)
    takes = DeviceTakes(dd, x, pos, all_ims, len(a_ims), tile_points=plan.tile_points)
    for got, nm in zip(takes.reference_arrays()[:4], ["m_take", "n_take", "p_take", "np_take"]):
        assert np.array_equal(got, g[nm]), nm
    params = pack_params(plan, [(torch.as_tensor(w, device=dev), torch.as_tensor(b, device=dev)) for w, b in layers])
    ev = ConstraintEvaluator(plan, takes, x, dd)
    ujets = ev.forward(params)
    prob = getattr(problems, cs["problem"])
    sp, _ = prob.init_params(**cs["pkw"])
    ap = {"static": {"problem": {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sp.items()}}, "trainable": {}}
    ujs = jet.ujs_constrained(ujets, x, prob.constraining_fn, ap)
    for j, uj in enumerate(ujs):
        e = common.rel_err(uj.cpu().numpy(), g[f"uj_fd_{j}"])
        assert e < 1e-5, (name, kernel, j, e)
    # value path (what FBPINN_model_jit returns): u and the window sums wp
    vplan = Plan(cs["layer_sizes"], JetSpec(((0, ()),), xd, ud), kernel=kernel)
    vev = ConstraintEvaluator(vplan, takes, x, dd)
    u = prob.constraining_fn(ap, x, vev.forward(params))
    assert common.rel_err(u.cpu().numpy(), g["u"]) < 1e-5
    assert common.rel_err(vev.dsum[:takes.q, 0].cpu().numpy(), g["wp"][:, 0]) < 1e-5
    us = vev.pair_values_reference_order()[:, 0].cpu().numpy()
    assert common.rel_err(us, g["us"][:, 0]) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_inference_api_matches_reference_golden(name):
    "analysis.FBPINN_model / FBPINN_solution (value-only twin of the hot path) on the reference-generated fixtures"
    from fbpinns_b200 import analysis, networks, domains
    from fbpinns_b200.constants import Constants
    g, cs, layers = _load(name)
    prob = getattr(problems, cs["problem"])
    sd, _ = decompositions.RectangularDecompositionND.init_params(**cs["dkw"])
    sp, tp = prob.init_params(**cs["pkw"])
    c = Constants(problem=prob, problem_init_kwargs=cs["pkw"], decomposition=decompositions.RectangularDecompositionND,
                  decomposition_init_kwargs=cs["dkw"], network=networks.FCN,
                  network_init_kwargs=dict(layer_sizes=cs["layer_sizes"]))
    all_params = {"static": {"problem": sp, "decomposition": sd},
                  "trainable": {"network": {"subdomain": {"layers": [(torch.as_tensor(w), torch.as_tensor(b)) for w, b in layers]}}}}
    u, wp, us = analysis.FBPINN_model(c, all_params, g["active_in"], torch.as_tensor(cs["x"]))
    assert common.rel_err(u.cpu().numpy(), g["u"]) < 1e-5
    assert common.rel_err(wp.cpu().numpy(), g["wp"]) < 1e-5
    assert common.rel_err(us.cpu().numpy(), g["us"]) < 1e-5
    assert torch.equal(analysis.FBPINN_solution(c, all_params, g["active_in"], torch.as_tensor(cs["x"])), u)


# ---------------------------------------------------------------------------------------------------------------
# refindex.npz: the reference's own batched inside tests and _get_update_inputs (rows A2-A4)

def _index_golden():
    return np.load(os.path.join(HERE, "golden", "refindex.npz"), allow_pickle=True)


@pytest.mark.parametrize("tag,name", [("ho", "ho1d_hardbc"), ("bg", "burgers2d")])
def test_oracle_inside_tests_match_reference(tag, name):
    g = _index_golden()
    cs = case_setup(name)
    assert np.array_equal(g[f"{tag}_x"], cs["x"])
    decomp = ref_takes.rectangular_init_params(**cs["dkw"])
    n_take, m_take, ims = ref_takes.inside_points(decomp, cs["x"])
    assert np.array_equal(n_take, g[f"{tag}_n_take"]) and np.array_equal(m_take, g[f"{tag}_m_take"])
    assert np.array_equal(ims, g[f"{tag}_inside_ims"])
    ips, d = ref_takes.inside_models(decomp, cs["x"], g[f"{tag}_models_sel"])
    assert np.array_equal(ips, g[f"{tag}_inside_ips"])
    assert abs(d - float(g[f"{tag}_d"])) <= 1e-6 * abs(float(g[f"{tag}_d"]))


def _ui_inputs(g):
    cs = case_setup("burgers2d")
    x1, x2, v2 = g["ui_x1"], g["ui_x2"], g["ui_v2"]
    assert np.array_equal(x1, cs["x"])
    cons = [[x1], [x2, v2]]
    xg = np.concatenate([x1, x2])
    offsets, fs = ref_takes.constraint_tables([len(x1), len(x2)])
    return cs, cons, xg, offsets, fs


def test_oracle_update_inputs_match_reference():
    "FBPINNTrainer._get_x_batch + _get_update_inputs of the reference vs the oracle, several 0/1/2 active masks"
    g = _index_golden()
    cs, cons, xg, offsets, fs = _ui_inputs(g)
    decomp = ref_takes.rectangular_init_params(**cs["dkw"])
    for t in range(int(g["ui_trials"])):
        ui = ref_takes.get_update_inputs(g[f"ui{t}_active_in"], decomp, xg, cons, fs, offsets)
        assert np.array_equal(ui["active"], g[f"ui{t}_active_out"])
        assert np.array_equal(ui["x_batch"], g[f"ui{t}_x_batch"])
        assert len(ui["active_ims"]) == int(g[f"ui{t}_n_active_params"]) and len(ui["fixed_ims"]) == int(g[f"ui{t}_n_fixed_params"])
        for ic, tk in enumerate(ui["takess"]):
            for got, nm in zip(tk[:4], ["m_take", "n_take", "p_take", "np_take"]):
                assert np.array_equal(got, g[f"ui{t}_c{ic}_{nm}"]), (t, ic, nm)
            assert tk[4] == int(g[f"ui{t}_c{ic}_npou"])
            for j, arr in enumerate(ui["constraints"][ic]):
                assert np.array_equal(arr, g[f"ui{t}_c{ic}_arr{j}"])


@pytest.mark.gpu
def test_device_update_inputs_match_reference():
    "the CUDA index construction (fbp_inside_count / fbp_takes_*) vs the reference's own _get_update_inputs output"
    from fbpinns_b200.engine import DeviceDecomposition
    from fbpinns_b200.trainers import get_update_inputs
    g = _index_golden()
    cs, cons, xg, offsets, fs = _ui_inputs(g)
    dev = torch.device("cuda:0")
    sd, _ = decompositions.RectangularDecompositionND.init_params(**cs["dkw"])
    dd = DeviceDecomposition(sd["subdomain"]["params"], sd["subdomain"]["pou"], dev)
    cons_d = [[torch.as_tensor(a, device=dev) for a in con] for con in cons]
    jets = [JetSpec(((0, ()), (0, (0,))), 2, 1), JetSpec(((0, ()),), 2, 1)]
    for t in range(int(g["ui_trials"])):
        inp = get_update_inputs(g[f"ui{t}_active_in"], None, dd, torch.as_tensor(xg, device=dev), cons_d, offsets, jets, [2, 16, 1])
        assert np.array_equal(inp.active, g[f"ui{t}_active_out"])
        assert np.array_equal(inp.x_batch.cpu().numpy(), g[f"ui{t}_x_batch"])
        assert len(inp.active_ims) == int(g[f"ui{t}_n_active_params"]) and len(inp.fixed_ims) == int(g[f"ui{t}_n_fixed_params"])
        for ic, tk in enumerate(inp.takess):
            for got, nm in zip(tk.reference_arrays()[:4], ["m_take", "n_take", "p_take", "np_take"]):
                assert np.array_equal(got, g[f"ui{t}_c{ic}_{nm}"]), (t, ic, nm)
            for j, arr in enumerate(inp.constraints[ic]):
                assert np.array_equal(arr.cpu().numpy(), g[f"ui{t}_c{ic}_arr{j}"])


@pytest.mark.parametrize("name", NAMES)
def test_oracle_parameter_gradients_match_finite_differences_of_reference(name):
    "reverse mode of the oracle vs central differences over parameters of L0 = sum R u, u from the reference's FBPINN_model"
    g, cs, layers = _load(name)
    dc, lc, takes, prob, ap, x, cf = _oracle_model(g, cs, layers)
    all_ims = g["all_ims"]
    full = [(torch.tensor(w, dtype=torch.float64, requires_grad=True), torch.tensor(b, dtype=torch.float64, requires_grad=True))
            for w, b in layers]
    ims = torch.as_tensor(all_ims, dtype=torch.long)
    lc = [(w[ims], b[ims]) for w, b in full]
    u = ref_model.fbpinn_model(dc, lc, x, takes, cf, ap)[0]
    L0 = (torch.as_tensor(g["grad_R"]) * u).sum()
    grads = torch.autograd.grad(L0, [t for wb in full for t in wb], allow_unused=True)
    for pick, fd in zip(g["grad_picks"], g["grad_fd"]):
        l, which = int(pick[0]), int(pick[1])
        gt = grads[2 * l + which]
        idx = tuple(int(v) for v in pick[2:2 + gt.dim()])
        got = float(gt[idx]) if gt is not None else 0.0
        assert abs(got - fd) <= 1e-6 * max(1.0, abs(fd)) + 1e-7 * float(np.abs(g["grad_fd"]).max()), (name, pick, got, fd)


def test_network_plugins_match_reference_source():
    """FCN / AdaptiveFCN / SIREN / AdaptiveSIREN / FourierFCN: the host mirrors (fbpinns_b200/networks.py, single point)
    and the oracle's pair-batched network_forward reproduce the outputs of the reference's own `network_fn`s
    (tests/golden/refnetworks.npz, made by make_golden_networks.py from fbpinns/networks.py:61-194 under the shim)."""
    from fbpinns_b200 import networks as N
    from oracle import ref_model
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refnetworks.npz"))
    x = torch.tensor(g["x"])
    for name, cls, n_extra in [("fcn", N.FCN, 0), ("adaptive_fcn", N.AdaptiveFCN, 1), ("siren", N.SIREN, 0),
                               ("adaptive_siren", N.AdaptiveSIREN, 2), ("fourier", N.FourierFCN, 0)]:
        layers = [tuple(torch.tensor(g[f"{name}_l{l}_{i}"]) for i in range(2 + n_extra)) for l in range(3)]
        st = {"network": {"subdomain": {"omega": torch.tensor(g["fourier_omega"])}}} if name == "fourier" else {}
        params = {"static": st, "trainable": {"network": {"subdomain": {"layers": layers}}}}
        y = torch.stack([cls.network_fn(params, xi) for xi in x]).numpy()
        assert np.abs(y - g[f"{name}_y"]).max() < 1e-13, name
        # oracle: every pair carries its own copy of the parameters
        s = x.shape[0]
        take = [tuple(t.unsqueeze(0).expand(s, *t.shape) for t in leaf) for leaf in layers]
        stt = {"omega": torch.tensor(g["fourier_omega"]).unsqueeze(0).expand(s, -1, -1)} if name == "fourier" else None
        yo = ref_model.network_forward(name, take, x, stt).numpy()
        assert np.abs(yo - g[f"{name}_y"]).max() < 1e-13, name
