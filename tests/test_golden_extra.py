"""Second batch of reference-generated fixtures (tests/golden/make_golden_extra.py, the REFERENCE'S OWN SOURCE under the
numpy jax-shim): MultilevelRectangularDecompositionND (npou = 2), HarmonicOscillator1D's two-constraint soft-BC loss,
HarmonicOscillator1DInverse with a trainable mu.  Pins the oracle on the CPU and, through the same fixtures, the CUDA
path on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_takes, ref_model
from fbpinns_b200 import problems, decompositions
from fbpinns_b200.jets import JetSpec
import common

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import cases_extra as CX  # noqa: E402

TAKES = ["m_take", "n_take", "p_take", "np_take"]


def _ml():
    g = np.load(os.path.join(HERE, "golden", "refextra_multilevel.npz"), allow_pickle=True)
    cs = CX.multilevel_setup()
    assert np.array_equal(g["x"], cs["x"]) and repr(cs["req"]) == str(g["req_repr"])
    layers = [(g[f"W{l}"], g[f"b{l}"]) for l in range(len(cs["layer_sizes"]) - 1)]
    return g, cs, layers


def test_multilevel_init_params_match_reference():
    g, cs, _ = _ml()
    ours, _ = decompositions.MultilevelRectangularDecompositionND.init_params(**cs["dkw"])
    orac = ref_takes.multilevel_init_params(**cs["dkw"])
    assert ours["m"] == orac["m"] == int(g["m"])
    for i in range(6):
        assert np.array_equal(np.asarray(ours["subdomain"]["params"][i]), g[f"static_{i}"]), i
        assert np.array_equal(orac["subdomain"]["params"][i], g[f"static_{i}"]), i
    assert np.array_equal(np.asarray(ours["subdomain"]["pou"]), g["pou"]) and np.array_equal(orac["subdomain"]["pou"], g["pou"])
    for nm in ("xmins0", "xmaxs0"):
        assert np.array_equal(np.asarray(ours[nm]), g[nm]) and np.array_equal(orac[nm], g[nm])


@pytest.mark.parametrize("trial", [0, 1])
def test_multilevel_get_inputs_and_model_match_reference(trial):
    "npou = 2: (point, pou) unique rows, per-level quotient, /npou — oracle vs the reference's FBPINN_model"
    g, cs, layers = _ml()
    t = f"t{trial}_"
    decomp = ref_takes.multilevel_init_params(**cs["dkw"])
    takes, all_ims, a_ims, f_ims, active = ref_takes.get_inputs(cs["x"], g[t + "active_in"], decomp)
    assert np.array_equal(active, g[t + "active_out"]) and np.array_equal(all_ims, g[t + "all_ims"])
    for got, nm in zip(takes[:4], TAKES):
        assert np.array_equal(got, g[t + nm]), nm
    assert takes[4] == int(g[t + "npou"]) == 2
    dt = torch.float64
    dc = ref_model.cut_decomp(ref_model.to_torch(decomp, dt), all_ims)
    lc = [(torch.as_tensor(w[all_ims], dtype=dt), torch.as_tensor(b[all_ims], dtype=dt)) for w, b in layers]
    x = torch.as_tensor(cs["x"], dtype=dt)
    u, wp, us, ws, us_raw = ref_model.fbpinn_model(dc, lc, x, takes, None, None)
    for got, nm in [(u, "u"), (wp, "wp"), (us, "us"), (ws, "ws"), (us_raw, "us_raw")]:
        assert np.allclose(got.numpy(), g[t + nm], rtol=1e-12, atol=1e-13), nm
    ujs = ref_model.fbpinn_forward(dc, lc, x, takes, ref_model.get_jmaps(cs["req"]), None, None)
    for j, uj in enumerate(ujs):
        assert common.rel_err(uj.numpy(), g[t + f"uj_fd_{j}"]) < 2e-6, j
    # reverse mode of the oracle against parameter finite differences of the reference's model
    R = torch.as_tensor(g[t + "grad_R"])
    leaves = [tuple(torch.tensor(a[all_ims], dtype=dt, requires_grad=True) for a in wb) for wb in layers]
    L0 = (R * ref_model.fbpinn_model(dc, leaves, x, takes, None, None)[0]).sum()
    grads = torch.autograd.grad(L0, [a for wb in leaves for a in wb])
    pos = -np.ones(decomp["m"], dtype=int)
    pos[all_ims] = np.arange(len(all_ims))
    for pick, fd in zip(g[t + "grad_picks"], g[t + "grad_fd"]):
        l, which, im = int(pick[0]), int(pick[1]), int(pick[2])
        gr = grads[2 * l + which]
        if pos[im] < 0:
            assert abs(fd) < 1e-9
            continue
        got = float(gr[pos[im], pick[3], pick[4]] if which == 0 else gr[pos[im], pick[3]])
        assert abs(got - fd) <= 1e-6 * max(1.0, float(gr.abs().max())), (pick, got, fd)


def _ho():
    g = np.load(os.path.join(HERE, "golden", "refextra_ho1d.npz"), allow_pickle=True)
    cs = CX.ho1d_setup()
    layers = [(g[f"W{l}"], g[f"b{l}"]) for l in range(len(cs["layer_sizes"]) - 1)]
    return g, cs, layers


class _Dom:
    def __init__(self, x):
        self.x = x

    def sample_interior(self, all_params, key, sampler, batch_shape):
        return torch.as_tensor(self.x)


@pytest.mark.parametrize("tag,prob", [("soft", problems.HarmonicOscillator1D), ("inv", problems.HarmonicOscillator1DInverse)])
def test_ho1d_problem_definitions_match_reference(tag, prob):
    "init_params, sample_constraints (incl. the inverse problem's data from exact_solution), required ujs"
    g, cs, _ = _ho()
    sp, tp = prob.init_params(**cs["pkw"])
    for k, v in sp.items():
        assert np.allclose(np.asarray(v, dtype=np.float64), g[f"{tag}_pstat_{k}"]), k
    for k, v in (tp or {}).items():
        assert np.allclose(np.asarray(v), g[f"{tag}_ptrain_{k}"])
    ap = {"static": {"problem": sp}, "trainable": {"problem": tp} if tp else {}}
    cons = prob.sample_constraints(ap, _Dom(cs["x_phys"]), None, "grid", ((len(cs["x_phys"]),),))
    assert repr([tuple(c_[-1]) for c_ in cons]) == str(g[f"{tag}_reqs_repr"])
    for ic, c_ in enumerate(cons):
        for j, a in enumerate(c_[:-1]):
            # the data constraint's u comes from exact_solution evaluated in float32 (as JAX would with x64 off); the
            # fixture holds the shim's float64 evaluation of the same reference code
            assert np.allclose(np.asarray(a, dtype=np.float64), g[f"{tag}_c{ic}_arr{j}"], rtol=1e-6, atol=2e-6), (ic, j)
    ue = prob.exact_solution(ap, torch.as_tensor(g[f"{tag}_exact_x"]))
    assert np.allclose(ue.numpy(), g[f"{tag}_exact_u"], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("tag,prob", [("soft", problems.HarmonicOscillator1D), ("inv", problems.HarmonicOscillator1DInverse)])
def test_ho1d_two_constraint_takes_loss_and_mu_gradient_match_reference(tag, prob):
    """per-constraint takes from the reference's _get_update_inputs, the two-constraint loss on finite-difference ujs of the
    reference's model (oracle ujs must agree with them), the empty boundary branch, dL/dmu of the inverse problem"""
    g, cs, layers = _ho()
    decomp = ref_takes.rectangular_init_params(**cs["dkw"])
    sp, tp = prob.init_params(**cs["pkw"])
    ap0 = {"static": {"problem": sp}, "trainable": {"problem": tp} if tp else {}}
    cons = prob.sample_constraints(ap0, _Dom(cs["x_phys"]), None, "grid", ((len(cs["x_phys"]),),))
    reqs = [c_[-1] for c_ in cons]
    # inputs = the arrays the reference's sample_constraints produced (ours agree to float32 round-off, checked above)
    cons_np = [[np.asarray(g[f"{tag}_c{ic}_arr{j}"], dtype=np.float32) for j in range(len(c_) - 1)] for ic, c_ in enumerate(cons)]
    xg = np.concatenate([c_[0] for c_ in cons_np])
    offsets, fs = ref_takes.constraint_tables([len(c_[0]) for c_ in cons_np])
    dt = torch.float64
    for trial in range(int(g[f"{tag}_n_trials"])):
        t = f"{tag}{trial}_"
        ui = ref_takes.get_update_inputs(g[t + "active_in"], decomp, xg, cons_np, fs, offsets)
        assert np.array_equal(ui["active"], g[t + "active_out"])
        dc = ref_model.cut_decomp(ref_model.to_torch(decomp, dt), ui["all_ims"])
        lc = [(torch.as_tensor(w[ui["all_ims"]], dtype=dt), torch.as_tensor(b[ui["all_ims"]], dtype=dt)) for w, b in layers]
        mu = torch.tensor(float(g[t + "mu"]), dtype=dt, requires_grad=True) if tp else None
        ap = {"static": {"problem": sp}, "trainable": {"problem": {"mu": mu}} if tp else {}}
        cons_fd, cons_or = [], []
        for ic, (tk, con) in enumerate(zip(ui["takess"], ui["constraints"])):
            for got, nm in zip(tk[:4], TAKES):
                assert np.array_equal(got, g[t + f"c{ic}_{nm}"]), (trial, ic, nm)
            assert np.array_equal(con[0], g[t + f"c{ic}_x"])
            x = torch.as_tensor(con[0], dtype=dt)
            extra = [torch.as_tensor(a, dtype=dt) for a in con[1:]]
            if len(x):
                ujs = ref_model.fbpinn_forward(dc, lc, x, tk, ref_model.get_jmaps(reqs[ic]), None, ap)
            else:
                ujs = [torch.zeros((0, 1), dtype=dt) for _ in reqs[ic]]
            for j, uj in enumerate(ujs):
                ref = g[t + f"c{ic}_uj_fd_{j}"]
                assert common.rel_err(uj.detach().numpy(), ref) < 5e-6, (trial, ic, j)
            cons_fd.append([x] + extra + [torch.as_tensor(g[t + f"c{ic}_uj_fd_{j}"]) for j in range(len(reqs[ic]))])
            cons_or.append([x] + extra + ujs)
        # the torch restatement of loss_fn on exactly the arrays the reference's loss_fn saw
        loss = prob.loss_fn(ap, cons_fd)
        ref_loss = float(g[t + "loss"])
        assert abs(float(loss) - ref_loss) <= 1e-10 * abs(ref_loss), (trial, float(loss), ref_loss)
        # ... and on the oracle's own ujs (differs only by the finite-difference error of the fixture)
        assert abs(float(prob.loss_fn(ap, cons_or)) - ref_loss) <= 2e-4 * abs(ref_loss)
        if tp:
            (dmu,) = torch.autograd.grad(loss, [mu])
            ref_d = float(g[t + "dloss_dmu_fd"])
            assert abs(float(dmu) - ref_d) <= 1e-6 * max(1.0, abs(ref_d)), (float(dmu), ref_d)


# ------------------------------------------------------------------------------------------------ CUDA path on the same fixtures

@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["generic", "auto"])
@pytest.mark.parametrize("trial", [0, 1])
def test_cuda_multilevel_matches_reference_golden(trial, kernel):
    from fbpinns_b200.engine import DeviceDecomposition, DeviceTakes, ConstraintEvaluator, Plan, pack_params
    from fbpinns_b200.trainers import active_set_algebra
    g, cs, layers = _ml()
    t = f"t{trial}_"
    dev = torch.device("cuda:0")
    sd, _ = decompositions.MultilevelRectangularDecompositionND.init_params(**cs["dkw"])
    dd = DeviceDecomposition(sd["subdomain"]["params"], sd["subdomain"]["pou"], dev)
    x = torch.as_tensor(cs["x"], device=dev)
    _, mc = dd.inside_count(x)
    act2, a_ims, f_ims, all_ims, pos = active_set_algebra(g[t + "active_in"], mc.cpu().numpy())
    assert np.array_equal(all_ims, g[t + "all_ims"]) and np.array_equal(act2, g[t + "active_out"])
    jet = JetSpec(cs["req"], 2, 1)
    plan = Plan(cs["layer_sizes"], jet, kernel=kernel)
    takes = DeviceTakes(dd, x, pos, all_ims, len(a_ims), tile_points=plan.tile_points)
    for got, nm in zip(takes.reference_arrays()[:4], TAKES):
        assert np.array_equal(got, g[t + nm]), nm
    assert takes.npou == 2
    params = pack_params(plan, [(torch.as_tensor(w, device=dev), torch.as_tensor(b, device=dev)) for w, b in layers])
    ev = ConstraintEvaluator(plan, takes, x, dd)
    ujs = jet.ujs_plain(ev.forward(params))
    for j, uj in enumerate(ujs):
        e = common.rel_err(uj.cpu().numpy(), g[t + f"uj_fd_{j}"])
        assert e < 1e-5, (trial, kernel, j, e)
    # reverse mode: gradient of sum_p R_p u(x_p) against the parameter finite differences of the reference's model
    ubar = torch.zeros((takes.n, ev.V), device=dev)
    ubar[:, jet.column(0, ())] = torch.as_tensor(g[t + "grad_R"][:, 0], dtype=torch.float32, device=dev)
    grads = torch.zeros((max(len(a_ims), 1), plan.P), device=dev)
    ev.backward(ubar, params, grads, accumulate=False)
    from fbpinns_b200.engine import unpack_params
    got = unpack_params(plan, grads)
    apos = {int(im): i for i, im in enumerate(a_ims)}
    scale = [max(float(gw.abs().max()), float(gb.abs().max())) for gw, gb in got]
    for pick, fd in zip(g[t + "grad_picks"], g[t + "grad_fd"]):
        l, which, im = int(pick[0]), int(pick[1]), int(pick[2])
        if im not in apos:
            continue                         # fixed / discarded subdomains carry no gradient
        v = float(got[l][0][apos[im], pick[3], pick[4]] if which == 0 else got[l][1][apos[im], pick[3]])
        assert abs(v - fd) <= 2e-5 * max(1.0, scale[l]), (pick, v, fd)
