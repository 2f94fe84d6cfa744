"""CPU tests: oracle self-checks and the kernel-maths prototype against the oracle's autograd."""
import numpy as np
import pytest
import torch

from oracle import ref_takes, ref_model, ref_adam
from fbpinns_b200 import configs
from fbpinns_b200.jets import JetSpec
import common
import proto_kernel_math as proto


def test_pair_ordering_matches_dense_nonzero():
    """The reference's only self-checking code (fbpinns/decompositions_base.py:87-128): n=10000 points in [0,2]^2,
    m=1000 boxes, pairs must be the row-major non-zeros of the dense inside mask for any batch size."""
    rng = np.random.default_rng(0)
    n, m = 10000, 1000
    x = rng.uniform(0, 2, (n, 2)).astype(np.float32)
    cc = rng.uniform(1, 3, (m, 2)).astype(np.float32)
    decomp = {"m": m, "xd": 2, "subdomain": {"params": [cc - np.float32(0.1), cc + np.float32(0.1)], "pou": np.zeros((m, 1), np.float32)}}
    dense = ref_takes.inside_mask(decomp, x, np.arange(m))
    nt, mt = np.nonzero(dense)
    for batch in [1, 9, 10, 128, n]:
        parts_n, parts_m = [], []
        for i0, msk in ref_takes._batched(decomp, x, np.arange(m), batch=batch):
            a, b = np.nonzero(msk)
            parts_n.append(a + i0)
            parts_m.append(b)
            if batch == 1 and i0 > 300:
                break
        if batch == 1:
            k = len(np.concatenate(parts_n))
            assert (np.concatenate(parts_n) == nt[:k]).all() and (np.concatenate(parts_m) == mt[:k]).all()
        else:
            assert (np.concatenate(parts_n) == nt).all() and (np.concatenate(parts_m) == mt).all()
    n_take, m_take, ims = ref_takes.inside_points(decomp, x)
    assert (n_take == nt).all() and (m_take == mt).all()
    assert (ims == np.nonzero(dense.any(0))[0]).all()
    ips, _ = ref_takes.inside_models(decomp, x, np.arange(m))
    assert (ips == np.nonzero(dense.any(1))[0]).all()


def test_get_jmaps_examples():
    # Burgers (fbpinns/problems.py:305-310): nodes (0,), (0,0), (1,) -> a 2nd-order chain in x and a 1st-order in t
    nodes, leaves, jac_is = ref_model.get_jmaps(((0, ()), (0, (0,)), (0, (1,)), (0, (0, 0))))
    assert nodes == (((0, 0), (0,), 0), ((1, 0), (0, 0), 1), ((0, 1), (1,), 1))
    assert leaves == ((2, (0, 0)), (3, (1,)))
    assert jac_is == ((0, 0, 0), (0, 1, 0), (1, 1, 0), (0, 2, 0))
    nodes, leaves, jac_is = ref_model.get_jmaps(((0, ()),))
    assert nodes == () and leaves == ((0, ()),) and jac_is == ((0, 0, 0),)
    from fbpinns_b200.jets import get_jmaps
    for req in [((0, ()), (0, (0,)), (0, (0, 0))), ((0, (0, 0)), (0, (1, 1)), (0, (2, 2))), ((0, (0, 1)), (1, (1,)))]:
        assert get_jmaps(req) == ref_model.get_jmaps(req)


def test_oracle_jets_against_closed_form():
    "single subdomain (flag = 0 -> window 1): u = tanh(w z + b) v + c, derivatives in closed form"
    decomp = ref_takes.rectangular_init_params([np.array([0.5])], [np.array([1.0])], (0.25, 2.0))
    dt = ref_model.to_torch(decomp, torch.float64)
    w0, b0, w1, b1 = 0.7, -0.2, 1.3, 0.4
    layers = [(torch.tensor([[[w0]]], dtype=torch.float64), torch.tensor([[b0]], dtype=torch.float64)),
              (torch.tensor([[[w1]]], dtype=torch.float64), torch.tensor([[b1]], dtype=torch.float64))]
    x = torch.linspace(0.1, 0.9, 7, dtype=torch.float64).reshape(-1, 1)
    takes = (np.zeros(7, np.int32), np.arange(7, dtype=np.int32), np.arange(7, dtype=np.int32), np.arange(7, dtype=np.int32), 1)
    jm = ref_model.get_jmaps(((0, ()), (0, (0,)), (0, (0, 0))))
    u, ux, uxx = [t.numpy() for t in ref_model.fbpinn_forward(dt, layers, x, takes, jm)]
    z = (x.numpy() - 0.5) / 0.5
    t = np.tanh(w0 * z + b0)
    g = 1 - t * t
    assert np.allclose(u, (w1 * t + b1) * 2.0 + 0.25, atol=1e-14)
    assert np.allclose(ux, 2.0 * w1 * g * w0 / 0.5, atol=1e-13)
    assert np.allclose(uxx, 2.0 * w1 * (-2 * t * g) * (w0 / 0.5) ** 2, atol=1e-12)


def test_adam_closed_form():
    "first step of Adam moves every parameter by -lr * sign(g) (up to eps); constant gradients keep doing so"
    p = [np.array([1.0, -2.0, 3.0], dtype=np.float64)]
    g = [np.array([0.5, -0.25, 2.0], dtype=np.float64)]
    st = ref_adam.adam_init(p)
    p1, st = ref_adam.adam_update(g, st, p, learning_rate=1e-3)
    assert np.allclose(p1[0], p[0] - 1e-3 * np.sign(g[0]), atol=1e-9)
    assert st["count"] == 1
    p2, st = ref_adam.adam_update(g, st, p1, learning_rate=1e-3)
    assert np.allclose(p2[0], p[0] - 2e-3 * np.sign(g[0]), atol=1e-9)
    # three steps with varying gradient against a straight transcription of the published formulas
    rng = np.random.default_rng(0)
    p = [rng.normal(size=5)]
    st = ref_adam.adam_init(p)
    m = v = np.zeros(5)
    pp = p[0].copy()
    for t in range(1, 4):
        gg = rng.normal(size=5)
        p, st = ref_adam.adam_update([gg], st, p)
        m = 0.9 * m + 0.1 * gg
        v = 0.999 * v + 0.001 * gg * gg
        pp = pp - 1e-3 * (m / (1 - 0.9 ** t)) / (np.sqrt(v / (1 - 0.999 ** t)) + 1e-8)
        assert np.allclose(p[0], pp, rtol=1e-12)


CASES = ["cfg1", "cfg2", "cfg3", "cfg5"]


@pytest.mark.parametrize("name", CASES)
def test_kernel_math_prototype_matches_oracle(name):
    """forward jets, row sums + quotient rule and the hand-derived reverse pass (what the CUDA kernels compute) agree
    with the oracle's nested-jvp + autograd formulation in float64."""
    small = dict(configs.SMALL[name])
    if name == "cfg3":
        small.update(n_sub=(3, 3), n_pts=(20, 20), line_scheduler=False)
    if name == "cfg5":
        small.update(n_sub=(3, 3), n_pts=(24, 24), layer_sizes=(2, 8, 8, 1))
    if name == "cfg1":
        small.update(n_pts=50)
    if name == "cfg2":
        small.update(n_sub=6, n_pts=50)
    k = common.make_case(configs.CONFIGS[name](**small), seed=1)
    ui = k.ui
    decomp = k.decomp_np
    ps = decomp["subdomain"]["params"]
    ss_all = np.concatenate([ps[0], ps[1], ps[4], ps[5]], axis=1).astype(np.float64)
    rng = np.random.default_rng(5)
    for ic, jet in enumerate(k.jets):
        takes = ui["takess"][ic]
        m_take, n_take, p_take, np_take, npou = takes
        if len(m_take) == 0:
            continue
        x = ui["constraints"][ic][0].astype(np.float64)
        im = ui["all_ims"][m_take]
        lt = [(w[im].astype(np.float64), b[im].astype(np.float64)) for w, b in k.layers]
        N, cache = proto.pair_forward(jet, x[n_take], ss_all[im], lt)
        wj = cache["w"]
        dsum = np.zeros((len(np_take), jet.C))
        np.add.at(dsum, p_take, wj)
        ujets = proto.reduce_forward(jet, N, dsum, takes, x.shape[0])
        ref = common.oracle_ujs(k, ic, torch.float64, constrained=False)
        for (iu, path), r in zip(jet.required_ujs, ref):
            got = ujets[:, jet.column(iu, path)]
            assert common.rel_err(got, r[:, 0]) < 1e-10, (name, ic, path)

        # reverse pass: L = sum_j <R_j, ujs_j>
        R = [rng.normal(size=(x.shape[0],)) for _ in jet.required_ujs]
        ubar = np.zeros((x.shape[0], jet.C * jet.ud))
        for (iu, path), r in zip(jet.required_ujs, R):
            ubar[:, jet.column(iu, path)] += r
        grow = proto.reduce_backward(jet, ubar, dsum, takes)
        m_active = len(ui["active_ims"])
        grads = proto.pair_backward(jet, lt, cache, grow[p_take], m_take, m_active)

        dtype = torch.float64
        decomp_cut = ref_model.cut_decomp(ref_model.to_torch(decomp, dtype), ui["all_ims"])
        lc = [(torch.tensor(w[ui["all_ims"]], dtype=dtype, requires_grad=True),
               torch.tensor(b[ui["all_ims"]], dtype=dtype, requires_grad=True)) for w, b in k.layers]
        ujs = ref_model.fbpinn_forward(decomp_cut, lc, torch.as_tensor(x), takes, k.jmapss[ic])
        L = sum((torch.as_tensor(r).reshape(-1, 1) * u).sum() for r, u in zip(R, ujs))
        gs = torch.autograd.grad(L, [t for wb in lc for t in wb])
        for l, (gW, gb) in enumerate(grads):
            assert common.rel_err(gW, gs[2 * l].numpy()[:m_active]) < 1e-9, (name, ic, l)
            assert common.rel_err(gb, gs[2 * l + 1].numpy()[:m_active]) < 1e-9, (name, ic, l)


def test_kernel_math_prototype_mixed_derivative():
    "mixed second derivative u_xy (generic kernel family only) through the same formulas"
    c = configs.cfg5_poisson(n_sub=(3, 3), n_pts=(16, 16), layer_sizes=(2, 8, 1))
    k = common.make_case(c, seed=2)
    jet = JetSpec(((0, (0, 1)), (0, (1, 1)), (0, ())), 2, 1)
    jm = ref_model.get_jmaps(jet.required_ujs)
    ui = k.ui
    takes = ui["takess"][0]
    m_take, n_take, p_take, np_take, npou = takes
    ps = k.decomp_np["subdomain"]["params"]
    ss_all = np.concatenate([ps[0], ps[1], ps[4], ps[5]], axis=1).astype(np.float64)
    x = ui["constraints"][0][0].astype(np.float64)
    im = ui["all_ims"][m_take]
    lt = [(w[im].astype(np.float64), b[im].astype(np.float64)) for w, b in k.layers]
    N, cache = proto.pair_forward(jet, x[n_take], ss_all[im], lt)
    dsum = np.zeros((len(np_take), jet.C))
    np.add.at(dsum, p_take, cache["w"])
    ujets = proto.reduce_forward(jet, N, dsum, takes, x.shape[0])
    dtype = torch.float64
    decomp_cut = ref_model.cut_decomp(ref_model.to_torch(k.decomp_np, dtype), ui["all_ims"])
    lc = [(torch.tensor(w[ui["all_ims"]], dtype=dtype, requires_grad=True),
           torch.tensor(b[ui["all_ims"]], dtype=dtype, requires_grad=True)) for w, b in k.layers]
    ujs = ref_model.fbpinn_forward(decomp_cut, lc, torch.as_tensor(x), takes, jm)
    for (iu, path), r in zip(jet.required_ujs, ujs):
        assert common.rel_err(ujets[:, jet.column(iu, path)], r.detach().numpy()[:, 0]) < 1e-10, path
    rng = np.random.default_rng(0)
    R = [rng.normal(size=(x.shape[0],)) for _ in jet.required_ujs]
    ubar = np.zeros((x.shape[0], jet.C))
    for (iu, path), r in zip(jet.required_ujs, R):
        ubar[:, jet.column(iu, path)] += r
    grow = proto.reduce_backward(jet, ubar, dsum, takes)
    grads = proto.pair_backward(jet, lt, cache, grow[p_take], m_take, k.m)
    L = sum((torch.as_tensor(r).reshape(-1, 1) * u).sum() for r, u in zip(R, ujs))
    gs = torch.autograd.grad(L, [t for wb in lc for t in wb])
    for l, (gW, gb) in enumerate(grads):
        assert common.rel_err(gW, gs[2 * l].numpy()) < 1e-9
        assert common.rel_err(gb, gs[2 * l + 1].numpy()) < 1e-9


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_activation_jet_formulas_match_autograd(kind):
    """proto_activation_math (the formulas of the generic kernels' activation variants: tanh, alpha*tanh(a/alpha), sin,
    c*sin(o*a) — the reference's FCN / AdaptiveFCN / SIREN / AdaptiveSIREN, fbpinns/networks.py:61-166) against torch
    autograd of the activation composed with a second-order Taylor path, forward jets and the full reverse pass
    (pre-activation cotangents and the activation parameters' gradients), including a mixed second derivative."""
    import proto_activation_math as pa
    jet = JetSpec(((0, (0, 0)), (0, (0, 1)), (0, (1, 1))), 2, 1)
    order = [len(p) for p in jet.comps]
    i1 = [jet.index[(p[0],)] if len(p) == 2 else 0 for p in jet.comps]
    i2 = [jet.index[(p[1],)] if len(p) == 2 else 0 for p in jet.comps]
    rng = np.random.default_rng(kind)
    B, C = 7, jet.C
    a = rng.standard_normal((B, C))
    hbar = rng.standard_normal((B, C))
    p = tuple(rng.uniform(0.5, 1.5, B) for _ in range(pa.N_EXTRA[kind]))

    def f(v, ps):
        if kind == 0:
            return torch.tanh(v)
        if kind == 1:
            return ps[0] * torch.tanh(v / ps[0])
        if kind == 2:
            return torch.sin(v)
        return ps[0] * torch.sin(ps[1] * v)

    h_ref = np.zeros((B, C))
    ab_ref = np.zeros((B, C))
    pb_ref = [np.zeros(B) for _ in p]
    for s in range(B):
        at = torch.tensor(a[s], dtype=torch.float64, requires_grad=True)
        pt = [torch.tensor(q[s], dtype=torch.float64, requires_grad=True) for q in p]
        tau = torch.zeros(2, dtype=torch.float64, requires_grad=True)
        path = at[0]
        for c, comp in enumerate(jet.comps):
            if len(comp) == 1:
                path = path + at[c] * tau[comp[0]]
            elif len(comp) == 2:
                path = path + (0.5 if comp[0] == comp[1] else 1.0) * at[c] * tau[comp[0]] * tau[comp[1]]
        hv = f(path, pt)
        g1, = torch.autograd.grad(hv, tau, create_graph=True)
        hs = []
        for c, comp in enumerate(jet.comps):
            if len(comp) == 0:
                hs.append(hv)
            elif len(comp) == 1:
                hs.append(g1[comp[0]])
            else:
                g2, = torch.autograd.grad(g1[comp[0]], tau, create_graph=True)
                hs.append(g2[comp[1]])
        hs = torch.stack(hs)
        h_ref[s] = hs.detach().numpy()
        L = (hs * torch.tensor(hbar[s])).sum()
        grads = torch.autograd.grad(L, [at] + pt)
        ab_ref[s] = grads[0].numpy()
        for i in range(len(p)):
            pb_ref[i][s] = grads[1 + i].item()
    h = pa.act_forward(kind, a, p, order, i1, i2)
    ab, pbar = pa.act_backward(kind, a, p, hbar, order, i1, i2)
    assert np.allclose(h, h_ref, rtol=1e-10, atol=1e-10)
    assert np.allclose(ab, ab_ref, rtol=1e-10, atol=1e-10)
    for got, ref in zip(pbar, pb_ref):
        assert np.allclose(got, ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("name", ["adaptive_fcn", "siren", "adaptive_siren", "fourier"])
def test_activation_variant_kernels_emulated_on_cpu_match_oracle(name, tmp_path):
    """csrc/fbp_generic_act.cu — the kernels' OWN SOURCE compiled as plain C++ (tests/tools/fbp_host_emu.h) and run one
    pair at a time on the CPU — against the oracle: per-pair numerator jets (including a mixed second derivative) and
    the gradients of every parameter leaf for a random cotangent.  This pins the kernels' indexing and arithmetic
    before any GPU time is spent on them (float32 kernels vs float64 oracle: 2e-5 relative)."""
    import ctypes as C
    import os
    import shutil
    import subprocess
    from fbpinns_b200 import _lib
    from fbpinns_b200.engine import Plan
    from fbpinns_b200 import networks as N
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / "libemu_act.so")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I" + os.path.join(root, "tests", "tools"),
                        "-I" + os.path.join(root, "fbpinns_b200", "csrc"), "-o", so,
                        os.path.join(root, "tests", "tools", "emu_generic_act.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    emu = C.CDLL(so)

    cls, oname, n_extra = {"adaptive_fcn": (N.AdaptiveFCN, "adaptive_fcn", 1), "siren": (N.SIREN, "siren", 0),
                           "adaptive_siren": (N.AdaptiveSIREN, "adaptive_siren", 2), "fourier": (N.FourierFCN, "fourier", 0)}[name]
    rng = np.random.default_rng(5)
    xd, m, nf = 2, 3, 2
    counts = [7, 4, 6]
    s = sum(counts)
    hidden = [xd, 5, 4, 1]
    sizes = ([2 * nf] + hidden[1:]) if name == "fourier" else hidden
    layers = []
    for fi, fo in zip(sizes[:-1], sizes[1:]):
        v = np.sqrt((6.0 if "siren" in name else 1.0) / fi)
        leaf = [rng.uniform(-v, v, (m, fo, fi)), rng.uniform(-v, v, (m, fo))] + [rng.uniform(0.6, 1.4, (m, fo)) for _ in range(n_extra)]
        layers.append(tuple(t.astype(np.float32) for t in leaf))
    ap = {"static": {}, "trainable": {"network": {"subdomain": {"layers": [tuple(torch.tensor(t) for t in leaf) for leaf in layers]}}}}
    omega = None
    if name == "fourier":
        omega = (2 * np.pi * (0.1 + 0.3 * rng.standard_normal((m, nf, xd)))).astype(np.float32)
        ap["static"]["network"] = {"subdomain": {"omega": torch.tensor(omega)}}
    activation, ksizes, klayers = N.kernel_layers(cls, ap, "cpu")
    req = ((0, ()), (0, (0,)), (0, (1,)), (0, (0, 0)), (0, (0, 1)), (0, (1, 1)))
    jet = JetSpec(req, xd, 1)
    plan = Plan(ksizes, jet, activation=activation)
    P, Cj = plan.P, jet.C
    # packed rows in the documented layout: per layer W, b, activation parameters
    params = np.concatenate([np.concatenate([t.reshape(m, -1).numpy() for t in leaf], axis=1) for leaf in klayers], axis=1).astype(np.float32)
    assert params.shape == (m, P)
    lo = rng.uniform(-1, 0, (m, xd)).astype(np.float32)
    hi = (lo + rng.uniform(0.5, 1.5, (m, xd))).astype(np.float32)
    sub_static = np.ascontiguousarray(np.concatenate([lo, hi, np.ones((m, 1)), np.full((m, 1), 0.1), np.full((m, 1), 1.3)], axis=1), dtype=np.float32)
    sub_of_pair = np.repeat(np.arange(m), counts).astype(np.int32)
    x = (lo[sub_of_pair] + (hi - lo)[sub_of_pair] * rng.uniform(0.05, 0.95, (s, xd))).astype(np.float32)
    ident = np.arange(s, dtype=np.int32)
    sub_ids = np.arange(m, dtype=np.int32)
    tv = _lib.TakesView()
    tv.n, tv.s, tv.q, tv.s_active = s, s, s, s
    tv.m_all, tv.m_active, tv.npou = m, m, 1
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    tv.d_sub_ids, tv.d_spair_point, tv.d_spair_row, tv.d_spair_sub = vp(sub_ids), vp(ident), vp(ident), vp(sub_of_pair)
    scratch = np.zeros(max(plan.scratch_per_pair, 1) * s, dtype=np.float32)
    pair_out = np.full((s, Cj), np.nan, dtype=np.float32)
    fp = lambda a: a.ctypes.data_as(C.c_void_p)
    emu.emu_act_forward(plan.handle, C.byref(tv), fp(x), fp(params), fp(sub_static), fp(pair_out), fp(scratch))
    grow = rng.standard_normal((s, Cj)).astype(np.float32)
    grads = np.zeros((m, P), dtype=np.float32)
    emu.emu_act_backward(plan.handle, C.byref(tv), fp(x), fp(params), fp(sub_static), fp(grow), fp(grads), fp(scratch))

    # ---- oracle, float64: every pair is a "point" of its own
    T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    lt = [tuple(T(t).requires_grad_(True) for t in leaf) for leaf in layers]
    sp = torch.as_tensor(sub_of_pair, dtype=torch.long)
    ps_take = [T(lo)[sp], T(hi)[sp], None, None, torch.ones((s, 1), dtype=torch.float64),
               torch.tensor([[0.1, 1.3]], dtype=torch.float64).expand(s, 2)]
    static_take = None if omega is None else {"omega": T(omega)[sp]}

    def u_fn(xb):
        return ref_model.model_inner(ps_take, [tuple(t[sp] for t in leaf) for leaf in lt], xb, oname, static_take)[0], ()
    ujs = ref_model.get_ujs(T(x), ref_model.get_jmaps(req), u_fn)
    L = 0.0
    for (iu, p), ref in zip(req, ujs):
        col = jet.column(iu, p)
        e = np.abs(pair_out[:, col] - ref[:, 0].detach().numpy()).max() / max(ref.detach().abs().max().item(), 1e-30)
        assert e < 2e-5, f"{name} d{p}: {e:.2e}"
        L = L + (T(grow[:, col]) * ref[:, 0]).sum()
    ref_grads = torch.autograd.grad(L, [t for leaf in lt for t in leaf], allow_unused=True)
    # split the emulated gradient rows by the packed layout
    off, got = 0, []
    for leaf in klayers:
        for t in leaf:
            nel = int(np.prod(t.shape[1:]))
            got.append(grads[:, off:off + nel].reshape((m,) + tuple(t.shape[1:])))
            off += nel
    if name == "fourier":
        assert np.all(got[0] == 0) and np.all(got[1] == 0)          # static feature layer: no gradient
        got = got[2:]
    for g, rg in zip(got, ref_grads):
        if rg is None:
            assert np.all(g == 0)
            continue
        e = np.abs(g - rg.numpy()).max() / max(np.abs(rg.numpy()).max(), 1e-30)
        assert e < 2e-5, f"{name} gradient leaf {tuple(g.shape)}: {e:.2e}"


def _emu_lib(tmp_path, name):
    "tests/tools/<name>.cpp (a CUDA source of csrc/ compiled as plain C++ through fbp_host_emu.h) -> ctypes library"
    import ctypes as C
    import os
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / f"lib{name}.so")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I" + os.path.join(root, "tests", "tools"),
                        "-I" + os.path.join(root, "fbpinns_b200", "csrc"), "-o", so,
                        os.path.join(root, "tests", "tools", name + ".cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return C.CDLL(so)


def _host_takes_view(takes, n, m_all, m_active):
    """The kernel-facing view (include/fbpinn_b200.h, fbp_takes_view) of reference-order takes, built with numpy the way
    csrc/fbp_takes.cu builds it on the device.  Returns (view, arrays kept alive)."""
    import ctypes as C
    from fbpinns_b200 import _lib
    m_take, n_take, p_take, np_take, npou = takes
    m_take, n_take, p_take, np_take = [np.ascontiguousarray(t, dtype=np.int32) for t in (m_take, n_take, p_take, np_take)]
    s, q = len(m_take), len(np_take)
    assert (np.diff(p_take) >= 0).all() and (np.diff(np_take) >= 0).all()
    perm = np.argsort(m_take, kind="stable")
    keep = dict(m_take=m_take, np_take=np_take, sub_ids=np.arange(m_all, dtype=np.int32),
                sub_off=np.concatenate([[0], np.cumsum(np.bincount(m_take, minlength=m_all))]).astype(np.int32),
                spair_point=n_take[perm].copy(), spair_row=p_take[perm].copy(), spair_sub=m_take[perm].copy(),
                pos=np.empty(s, dtype=np.int32),
                row_off=np.searchsorted(p_take, np.arange(q + 1)).astype(np.int32),
                pt_row_off=np.searchsorted(np_take, np.arange(n + 1)).astype(np.int32))
    keep["pos"][perm] = np.arange(s, dtype=np.int32)
    tv = _lib.TakesView()
    tv.n, tv.s, tv.q, tv.s_active = n, s, q, int(keep["sub_off"][m_active])
    tv.m_all, tv.m_active, tv.npou = m_all, m_active, int(npou)
    for field, key in [("d_m_take", "m_take"), ("d_np_take", "np_take"), ("d_sub_ids", "sub_ids"), ("d_sub_off", "sub_off"),
                       ("d_spair_point", "spair_point"), ("d_spair_row", "spair_row"), ("d_spair_sub", "spair_sub"),
                       ("d_pos", "pos"), ("d_row_off", "row_off"), ("d_pt_row_off", "pt_row_off")]:
        setattr(tv, field, keep[key].ctypes.data_as(C.c_void_p))
    return tv, keep


def test_generic_and_streaming_kernels_emulated_on_cpu_match_oracle(tmp_path):
    """The CUDA sources of the generic family (csrc/fbp_generic.cu) and of the streaming kernels (csrc/fbp_reduce.cu:
    window sums, row sums, segment-sum + quotient rule and its transpose, Adam) compiled as plain C++ and run on the CPU,
    chained exactly like a training step: ujs and parameter gradients against the oracle on a reduced Burgers case
    (fixed + active subdomains, mixed pair counts), one Adam step against the restated optax formula."""
    import ctypes as C
    from fbpinns_b200 import configs
    from fbpinns_b200.engine import Plan
    from oracle import ref_adam
    import common
    gen, red = _emu_lib(tmp_path, "emu_generic"), _emu_lib(tmp_path, "emu_reduce")
    c = configs.cfg3_burgers(n_sub=(4, 3), n_pts=(17, 13), layer_sizes=(2, 6, 5, 1), line_scheduler=False)
    active = np.ones(12, dtype=int)
    active[[2, 7]] = 2                                  # two fixed subdomains: forward only
    k = common.make_case(c, seed=3, active=active)
    ui = k.ui
    jet = k.jets[0]
    plan = Plan(k.layer_sizes, jet, kernel="generic")
    P, Cj = plan.P, jet.C
    all_ims, m_active = np.asarray(ui["all_ims"]), len(ui["active_ims"])
    x = np.ascontiguousarray(ui["constraints"][0][0], dtype=np.float32)
    n = x.shape[0]
    tv, keep = _host_takes_view(ui["takess"][0], n, len(all_ims), m_active)
    # kernels index the decomposition records and parameter rows by GLOBAL subdomain index through d_sub_ids
    keep["sub_ids"][:] = all_ims
    d = k.all_params["static"]["decomposition"]["subdomain"]["params"]
    f32 = lambda t: np.asarray(t, dtype=np.float32)
    sub_static = np.ascontiguousarray(np.concatenate([f32(d[0]), f32(d[1]), f32(d[4]), f32(d[5])], axis=1))
    params = np.ascontiguousarray(np.concatenate([np.concatenate([w.reshape(k.m, -1), b], axis=1) for w, b in k.layers], axis=1), dtype=np.float32)
    assert params.shape == (k.m, P) and sub_static.shape == (k.m, 2 * k.xd + 3)
    fp = lambda a: a.ctypes.data_as(C.c_void_p)
    s, q = tv.s, tv.q
    dsum = np.zeros((q, Cj), np.float32)
    pair_out = np.full((s, Cj), np.nan, np.float32)
    scratch = np.zeros(max(plan.scratch_per_pair, 1) * s, np.float32)
    ujets = np.full((n, Cj), np.nan, np.float32)
    red.emu_window_sums(plan.handle, C.byref(tv), fp(x), fp(sub_static), fp(dsum))
    gen.emu_generic_forward(plan.handle, C.byref(tv), fp(x), fp(params), fp(sub_static), fp(pair_out), fp(scratch))
    red.emu_reduce_forward(plan.handle, C.byref(tv), fp(pair_out), 0, fp(dsum), None, fp(ujets))
    ref = common.oracle_ujs(k, 0, torch.float64, constrained=False)
    for (iu, p), r in zip(jet.required_ujs, ref):
        e = common.rel_err(ujets[:, jet.column(iu, p)], r[:, 0])
        assert e < 2e-5, f"emulated ujs d{p}: {e:.2e}"
    # the split used by the multi-GPU halo step gives the same jets
    nsum = np.zeros((q, Cj), np.float32)
    ujets2 = np.zeros_like(ujets)
    red.emu_row_sums(plan.handle, C.byref(tv), fp(pair_out), fp(nsum))
    red.emu_reduce_forward(plan.handle, C.byref(tv), fp(nsum), 1, fp(dsum), None, fp(ujets2))
    assert np.array_equal(ujets, ujets2)

    # ---- reverse: cotangent of the jets -> rows -> parameter gradients of the ACTIVE subdomains
    rng = np.random.default_rng(0)
    ubar = rng.standard_normal((n, Cj)).astype(np.float32)
    grow = np.zeros((q, Cj), np.float32)
    grads = np.zeros((m_active, P), np.float32)
    red.emu_reduce_backward(plan.handle, C.byref(tv), fp(ubar), fp(dsum), None, fp(grow))
    gen.emu_generic_backward(plan.handle, C.byref(tv), fp(x), fp(params), fp(sub_static), fp(grow), fp(grads), fp(scratch))
    decomp_cut = ref_model.cut_decomp(ref_model.to_torch(k.decomp_np, torch.float64), all_ims)
    lc = [tuple(torch.tensor(t[all_ims], dtype=torch.float64, requires_grad=True) for t in leaf) for leaf in k.layers]
    ujs = ref_model.fbpinn_forward(decomp_cut, lc, torch.as_tensor(x, dtype=torch.float64), ui["takess"][0], k.jmapss[0])
    L = sum((torch.tensor(ubar[:, jet.column(iu, p)], dtype=torch.float64) * u[:, 0]).sum() for (iu, p), u in zip(jet.required_ujs, ujs))
    assert {jet.column(iu, p) for iu, p in jet.required_ujs} == set(range(Cj))
    rg = torch.autograd.grad(L, [t for leaf in lc for t in leaf])
    off = 0
    for t, g in zip([t for leaf in lc for t in leaf], rg):
        nel = int(np.prod(t.shape[1:]))
        got = grads[:, off:off + nel].reshape((m_active,) + tuple(t.shape[1:]))
        e = common.rel_err(got, g.numpy()[:m_active])
        assert e < 2e-5, f"emulated gradient {tuple(t.shape)}: {e:.2e}"
        assert float(g[m_active:].abs().max()) > 0            # the fixed subdomains do have a gradient the kernels skip
        off += nel

    # ---- the component-count instances of the two quotient-rule kernels (what the GPU launches for ud = 1, C <= 7) against
    #      the general kernels: same operations in the same order -> bit-identical, with and without an affine operator
    aff = rng.standard_normal((n, 2 * Cj)).astype(np.float32)
    for a_ in (None, aff):
        ap = None if a_ is None else fp(a_)
        u_gen, u_ct = np.full((n, Cj), np.nan, np.float32), np.full((n, Cj), np.nan, np.float32)
        red.emu_reduce_forward(plan.handle, C.byref(tv), fp(pair_out), 0, fp(dsum), ap, fp(u_gen))
        assert red.emu_reduce_forward_ct(plan.handle, C.byref(tv), fp(pair_out), 0, fp(dsum), ap, fp(u_ct), Cj) == 0
        assert np.array_equal(u_gen, u_ct)
        assert red.emu_reduce_forward_ct(plan.handle, C.byref(tv), fp(nsum), 1, fp(dsum), ap, fp(u_ct), Cj) == 0
        assert np.array_equal(u_gen, u_ct)
        g_gen, g_ct = np.full((q, Cj), np.nan, np.float32), np.full((q, Cj), np.nan, np.float32)
        red.emu_reduce_backward(plan.handle, C.byref(tv), fp(ubar), fp(dsum), ap, fp(g_gen))
        assert red.emu_reduce_backward_ct(plan.handle, C.byref(tv), fp(ubar), fp(dsum), ap, fp(g_ct), Cj) == 0
        assert np.array_equal(g_gen, g_ct)

    # ---- Adam on the active rows
    pr, mu, nu = params.copy(), np.zeros_like(params), np.zeros_like(params)
    rows = np.ascontiguousarray(all_ims[:m_active], dtype=np.int32)
    count = np.zeros(1, np.int32)
    red.emu_adam(fp(pr), fp(mu), fp(nu), fp(grads), fp(rows), C.c_int64(m_active), C.c_int64(P), fp(count),
                 C.c_float(1e-3), C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8), C.c_float(0.0))
    ref_p, _ = ref_adam.adam_update([grads], ref_adam.adam_init([params[rows]]), [params[rows]], learning_rate=1e-3)
    assert np.allclose(pr[rows], ref_p[0], rtol=2e-6, atol=1e-7)
    untouched = np.setdiff1d(np.arange(k.m), rows)
    assert np.array_equal(pr[untouched], params[untouched])


# --------------------------------------------------------------------------------------------------- Adam vs an independent implementation

def _torch_adam_run(p0_list, grads_per_step, join_step, dtype, lr=1e-3):
    """torch.optim.Adam (an implementation independent of the oracle's restatement: m / (sqrt(v)/sqrt(bc2) + eps) * lr/bc1)
    driven like the reference drives optax: ONE global step count; a parameter that joins the active set at `join_step`
    starts with zero moments but the global count (fbpinns/trainers.py:51-60, 528-531, 640)."""
    ps = [torch.tensor(p, dtype=dtype, requires_grad=True) for p in p0_list]
    opts = [torch.optim.Adam([p], lr=lr, betas=(0.9, 0.999), eps=1e-8) for p in ps]
    for it, grads in enumerate(grads_per_step):
        for i, (p, g, opt) in enumerate(zip(ps, grads, opts)):
            if it < join_step[i]:
                continue
            if it == join_step[i] and it > 0:       # late joiner: zero moments, global count
                opt.state[p]["step"] = torch.tensor(float(it))
                opt.state[p]["exp_avg"] = torch.zeros_like(p)
                opt.state[p]["exp_avg_sq"] = torch.zeros_like(p)
            p.grad = torch.as_tensor(g, dtype=dtype)
            opt.step()
    return [p.detach().numpy() for p in ps]


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 3e-6)])
def test_oracle_adam_matches_torch_optim_adam_50_steps(dtype, tol):
    """A10 pin: optax is absent from the reference tree, so the oracle's Adam is cross-checked against an independent
    implementation of the same published update (torch.optim.Adam, CPU) over 50 steps with gradients spanning six
    decades, including a subdomain that joins the active set late."""
    from oracle import ref_adam
    rng = np.random.default_rng(5)
    npdt = np.float64 if dtype == torch.float64 else np.float32
    p0 = [rng.normal(size=(4, 9)).astype(npdt), rng.normal(size=(3,)).astype(npdt)]
    join = [0, 20]
    steps = 50
    grads = [[(rng.normal(size=p.shape) * 10.0 ** rng.integers(-3, 3)).astype(npdt) for p in p0] for _ in range(steps)]
    want = _torch_adam_run(p0, grads, join, dtype)
    # oracle: the late joiner is held out of the tree, then merged with zero moments and the global count
    params, st = [p0[0].copy()], ref_adam.adam_init([p0[0]])
    for it in range(steps):
        if it == join[1]:
            params.append(p0[1].copy())
            st = dict(count=st["count"], mu=st["mu"] + [np.zeros_like(p0[1])], nu=st["nu"] + [np.zeros_like(p0[1])])
        params, st = ref_adam.adam_update(grads[it][:len(params)], st, params, learning_rate=1e-3)
    assert int(st["count"]) == steps
    for a, b in zip(params, want):
        assert np.max(np.abs(a - b)) <= tol * max(1.0, np.max(np.abs(b))), np.max(np.abs(a - b))
