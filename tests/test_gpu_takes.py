"""GPU parity: device index construction (fbp_inside_count / fbp_takes_*) is BIT-EXACT with the oracle's
restatement of inside_points / inside_models / get_inputs / _get_update_inputs."""
import numpy as np
import pytest
import torch

from oracle import ref_takes
from fbpinns_b200 import configs
from fbpinns_b200.engine import DeviceDecomposition, DeviceTakes, nonzero_i32
from fbpinns_b200.trainers import active_set_algebra
import common

pytestmark = pytest.mark.gpu


def _check_case(k, inp):
    ui = k.ui
    assert (inp.active == ui["active"]).all()
    assert (inp.active_ims == ui["active_ims"]).all()
    assert (inp.fixed_ims == ui["fixed_ims"]).all()
    assert (inp.all_ims == ui["all_ims"]).all()
    assert (inp.training_ips.cpu().numpy() == ui["training_ips"]).all()
    assert np.array_equal(inp.x_batch.cpu().numpy(), ui["x_batch"])
    for ic, t in enumerate(inp.takess):
        m_take, n_take, p_take, np_take, npou = t.reference_arrays()
        r = ui["takess"][ic]
        assert m_take.dtype == np.int32
        assert np.array_equal(m_take, r[0]), f"m_take constraint {ic}"
        assert np.array_equal(n_take, r[1]), f"n_take constraint {ic}"
        assert np.array_equal(p_take, r[2]), f"p_take constraint {ic}"
        assert np.array_equal(np_take, r[3]), f"np_take constraint {ic}"
        assert npou == r[4]
        for a, b in zip(inp.constraints[ic], ui["constraints"][ic]):
            assert np.array_equal(a.cpu().numpy(), b)
        # subdomain-sorted view: a stable permutation of the reference order
        pos = t.pos.cpu().numpy()
        assert np.array_equal(np.sort(pos), np.arange(t.s))
        assert np.array_equal(t.spair_sub.cpu().numpy()[pos], m_take)
        assert np.array_equal(t.spair_point.cpu().numpy()[pos], n_take)
        assert np.array_equal(t.spair_row.cpu().numpy()[pos], p_take)
        ssub = t.spair_sub.cpu().numpy()
        assert (np.diff(ssub) >= 0).all()
        spt = t.spair_point.cpu().numpy()
        same = np.diff(ssub) == 0
        assert (np.diff(spt)[same] > 0).all()
        so = t.sub_off.cpu().numpy()
        assert so[0] == 0 and so[-1] == t.s
        assert np.array_equal(np.diff(so), np.bincount(m_take, minlength=t.m_all))
        ro = t.row_off.cpu().numpy()
        assert np.array_equal(np.diff(ro), np.bincount(p_take, minlength=t.q))
        items = t.items_host
        assert items[:, 2].sum() == t.s
        assert (items[:, 2] > 0).all()


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_takes_all_active(name):
    import gpu_common
    k = common.make_case(configs.CONFIGS[name](**configs.SMALL[name]), seed=0)
    _, inp, _ = gpu_common.device_case(k)
    _check_case(k, inp)


def test_takes_scheduler_states():
    "active sets with fixed / inactive subdomains taken from the line scheduler (cfg 3), incl. models without points"
    import gpu_common
    from fbpinns_b200.schedulers import LineSchedulerRectangularND
    c = configs.cfg3_burgers(n_sub=(5, 5), n_pts=(40, 40), n_steps=40)
    k0 = common.make_case(c, seed=0)
    sched = LineSchedulerRectangularND(k0.all_params, 40, point=[0.], iaxis=0)
    states = [a.copy() for a in sched if a is not None]
    assert len(states) >= 3
    for active in states:
        k = common.make_case(c, seed=0, active=active)
        _, inp, _ = gpu_common.device_case(k)
        _check_case(k, inp)


def test_takes_multilevel_pou():
    "npou = 2 (MultilevelRectangularDecompositionND): rows are unique (point, pou) pairs"
    import gpu_common
    from fbpinns_b200 import decompositions
    from fbpinns_b200.constants import get_subdomain_ws
    xs1, xs2 = [np.linspace(0, 1, 3)], [np.linspace(0, 1, 7)]
    c = configs.cfg1_harmonic_oscillator(n_pts=60)
    c.decomposition = decompositions.MultilevelRectangularDecompositionND
    c.decomposition_init_kwargs = dict(subdomain_xss=[xs1, xs2],
                                       subdomain_wss=[get_subdomain_ws(xs1, 2.5), get_subdomain_ws(xs2, 2.5)], unnorm=(0., 1.))
    k = common.make_case(c, seed=0, multilevel=True)
    assert k.ui["takess"][0][4] == 2
    _, inp, _ = gpu_common.device_case(k)
    _check_case(k, inp)


def test_takes_random_boxes_and_uncovered_points():
    "arbitrary (non-grid) boxes, points outside every box, empty input — the reference's self-test geometry"
    rng = np.random.default_rng(0)
    n, m = 5000, 300
    x = rng.uniform(0, 2, (n, 2)).astype(np.float32)
    cc = rng.uniform(1, 3, (m, 2)).astype(np.float32)
    lo, hi = cc - np.float32(0.1), cc + np.float32(0.1)
    ones = np.ones((m, 1), np.float32)
    params = [lo, hi, ones, ones, ones, np.concatenate([0 * ones, ones], 1)]
    decomp = {"m": m, "xd": 2, "subdomain": {"params": params, "pou": 0 * ones}}
    dev = torch.device("cuda:0")
    dd = DeviceDecomposition(params, 0 * ones, dev)
    xd_ = torch.as_tensor(x, device=dev)
    pt, mc = dd.inside_count(xd_)
    dense = ref_takes.inside_mask(decomp, x, np.arange(m))
    assert np.array_equal(pt.cpu().numpy(), dense.sum(1))
    assert np.array_equal(mc.cpu().numpy(), dense.sum(0))
    assert np.array_equal(nonzero_i32(pt).cpu().numpy(), np.nonzero(dense.any(1))[0])
    sel = np.sort(rng.choice(m, 40, replace=False)).astype(np.int32)
    pt2, mc2 = dd.inside_count(xd_, models=torch.as_tensor(sel, device=dev))
    assert np.array_equal(pt2.cpu().numpy(), dense[:, sel].sum(1))
    assert np.array_equal(mc2.cpu().numpy(), dense[:, sel].sum(0))
    takes_ref, all_ims, a_ims, f_ims, act = ref_takes.get_inputs(x, np.ones(m, int), decomp)
    _, a2, f2, all2, pos = active_set_algebra(np.ones(m, int), mc.cpu().numpy())
    assert np.array_equal(all2, all_ims)
    t = DeviceTakes(dd, xd_, pos, all2, len(a2))
    got = t.reference_arrays()
    for a, b in zip(got[:4], takes_ref[:4]):
        assert np.array_equal(a, b)
    # empty input
    t0 = DeviceTakes(dd, xd_[:0].contiguous(), pos, all2, len(a2))
    assert t0.s == 0 and t0.q == 0 and t0.n == 0
