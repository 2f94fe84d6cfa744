"""CPU tests of the host-side logic (no GPU, no compute calls into the library)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import ref_takes, ref_model
from fbpinns_b200 import configs, decompositions, schedulers, _lib
from fbpinns_b200.engine import build_work_items, launch_records, one_item_per_subdomain
from fbpinns_b200.jets import JetSpec, get_ujs
from fbpinns_b200.trainers import active_set_algebra
from fbpinns_b200.constants import Constants, get_subdomain_ws

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    "the C-ABI library loads and exports every symbol include/fbpinn_b200.h declares (no compute calls)"
    hdr = open(os.path.join(ROOT, "include", "fbpinn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fbp_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.fbp_version.restype = ctypes.c_int
    assert lib.fbp_version() >= 100


def test_plan_validation_without_gpu():
    "plan creation is pure host code: closure / range errors are reported through fbp_last_error"
    from fbpinns_b200.engine import Plan
    p = Plan([2, 32, 32, 1], JetSpec(((0, (0, 0)), (0, (1, 1))), 2, 1))
    assert p.P == 2 * 32 + 32 + 32 * 32 + 32 + 32 + 1 == 1185
    with pytest.raises(NotImplementedError):
        JetSpec(((0, (0, 0, 0)),), 2, 1)
    with pytest.raises(_lib.FbpError):
        Plan([3, 8, 1], JetSpec(((0, ()),), 2, 1))          # layer_sizes[0] != xd


def test_no_cpu_fallback():
    from fbpinns_b200.engine import pack_params, Plan
    p = Plan([1, 4, 1], JetSpec(((0, ()),), 1, 1))
    with pytest.raises(_lib.FbpError):
        _lib.ptr(torch.zeros(3))


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_decomposition_init_matches_oracle_bitwise(name):
    c = configs.CONFIGS[name](**configs.SMALL[name])
    sd, _ = c.decomposition.init_params(**c.decomposition_init_kwargs)
    ref = ref_takes.rectangular_init_params(**c.decomposition_init_kwargs)
    assert sd["m"] == ref["m"] and sd["xd"] == ref["xd"]
    for a, b in zip(sd["subdomain"]["params"], ref["subdomain"]["params"]):
        assert a.dtype == torch.float32 and np.array_equal(a.numpy(), b)
    assert np.array_equal(sd["subdomain"]["pou"].numpy(), ref["subdomain"]["pou"])
    assert np.array_equal(sd["xmins0"], ref["xmins0"]) and np.array_equal(sd["xmaxs0"], ref["xmaxs0"])


def test_multilevel_and_non_overlapping():
    xs1, xs2 = [np.linspace(0, 1, 3), np.linspace(0, 2, 4)], [np.linspace(0, 1, 5), np.linspace(0, 2, 6)]
    kw = dict(subdomain_xss=[xs1, xs2], subdomain_wss=[get_subdomain_ws(xs1, 2.2), get_subdomain_ws(xs2, 2.2)], unnorm=(0., 2.))
    sd, _ = decompositions.MultilevelRectangularDecompositionND.init_params(**kw)
    ref = ref_takes.multilevel_init_params(**kw)
    assert sd["m"] == 12 + 30 == ref["m"]
    for a, b in zip(sd["subdomain"]["params"], ref["subdomain"]["params"]):
        assert np.array_equal(a.numpy(), b)
    assert np.array_equal(sd["subdomain"]["pou"].numpy(), ref["subdomain"]["pou"])
    with pytest.raises(ValueError):
        decompositions.RectangularDecompositionND.init_params([np.linspace(0, 1, 5)], [0.1 * np.ones(5)], (0., 1.))


def test_active_set_algebra_matches_oracle_get_inputs():
    rng = np.random.default_rng(0)
    c = configs.cfg3_burgers(n_sub=(5, 4), n_pts=(30, 30))
    decomp = ref_takes.rectangular_init_params(**c.decomposition_init_kwargs)
    m = decomp["m"]
    x = rng.uniform([-1, 0], [0.2, 0.5], size=(400, 2)).astype(np.float32)      # leaves some models without points
    for trial in range(5):
        active = rng.integers(0, 3, size=m)
        takes, all_ims, a_ims, f_ims, act = ref_takes.get_inputs(x, active, decomp)
        counts = ref_takes.inside_mask(decomp, x, np.arange(m)).sum(0)
        act2, a2, f2, all2, pos = active_set_algebra(active, counts)
        assert np.array_equal(act, act2) and np.array_equal(a_ims, a2) and np.array_equal(f_ims, f2)
        assert np.array_equal(all_ims, all2)
        assert (pos[all2] == np.arange(len(all2))).all() and (np.delete(pos, all2) == -1).all()


def test_schedulers_against_golden():
    "golden vectors produced by the reference's own fbpinns/schedulers.py (tests/golden/make_golden.py)"
    path = os.path.join(ROOT, "tests", "golden", "schedulers.npz")
    g = np.load(path, allow_pickle=True)
    c = configs.cfg3_burgers(n_sub=(6, 5), n_pts=(10, 10))
    sd, _ = c.decomposition.init_params(**c.decomposition_init_kwargs)
    ap = {"static": {"decomposition": sd}}
    runs = {
        "line": schedulers.LineSchedulerRectangularND(ap, 50, point=[0.], iaxis=0),
        "point": schedulers.PointSchedulerRectangularND(ap, 37, point=[0.3, 0.1]),
        "all": schedulers.AllActiveSchedulerND(ap, 5),
    }
    for name, sch in runs.items():
        steps = g[name + "_steps"]
        states = g[name + "_states"]
        got_steps, got_states = [], []
        for i, a in enumerate(sch):
            if a is not None:
                got_steps.append(i)
                got_states.append(np.array(a).copy())
        assert np.array_equal(np.array(got_steps), steps), name
        assert np.array_equal(np.array(got_states), states), name
    c3 = configs.cfg4_wave3d(n_sub=(3, 4, 5), n_pts=(4, 4, 4))
    sd3, _ = c3.decomposition.init_params(**c3.decomposition_init_kwargs)
    sch = schedulers.PlaneSchedulerRectangularND({"static": {"decomposition": sd3}}, 30, point=[0.], iaxes=[0, 1])
    got = [(i, np.array(a).copy()) for i, a in enumerate(sch) if a is not None]
    assert np.array_equal(np.array([i for i, _ in got]), g["plane_steps"])
    assert np.array_equal(np.array([a for _, a in got]), g["plane_states"])


def test_constraining_jets_by_local_taylor_model():
    "ujs of constraining_fn(x, u(x)) from the jets of u: exact against nested jvp through an analytic u"
    from fbpinns_b200.problems import BurgersEquation2D, WaveEquationConstantVelocity3D
    torch.manual_seed(0)
    for prob, xd, req in [(BurgersEquation2D, 2, ((0, ()), (0, (0,)), (0, (1,)), (0, (0, 0)))),
                          (WaveEquationConstantVelocity3D, 3, ((0, (0, 0)), (0, (1, 1)), (0, (2, 2))))]:
        sp, _ = prob.init_params()
        ap = {"static": {"problem": {k: (v.double() if torch.is_tensor(v) else v) for k, v in sp.items()}}, "trainable": {}}
        x = torch.rand(17, xd, dtype=torch.float64) * 0.8 + 0.1
        A = torch.randn(xd, dtype=torch.float64)

        def u_of(x):
            return torch.sin(x @ A).unsqueeze(1) + (x ** 2).sum(1, keepdim=True)
        jet = JetSpec(req, xd, 1)
        # analytic jets of u
        cols = []
        s_, c_ = torch.sin(x @ A), torch.cos(x @ A)
        for p in jet.comps:
            if len(p) == 0:
                cols.append(s_ + (x ** 2).sum(1))
            elif len(p) == 1:
                cols.append(c_ * A[p[0]] + 2 * x[:, p[0]])
            else:
                cols.append(-s_ * A[p[0]] * A[p[1]] + (2.0 if p[0] == p[1] else 0.0))
        ujets = torch.stack(cols, 1)
        got = jet.ujs_constrained(ujets, x, prob.constraining_fn, ap)
        ref = get_ujs(x, jet.jmaps, lambda xx: (prob.constraining_fn(ap, xx, u_of(xx)), ()))
        for a, b in zip(got, ref):
            assert torch.allclose(a, b, rtol=1e-10, atol=1e-10)


def test_work_items_cover_all_pairs():
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 3000, size=50)
    sub_off = np.concatenate([[0], np.cumsum(counts)])
    for tile, target, split in [(128, 1184, "equal"), (64, 10, "equal"), (128, 1, "equal"), (128, 1184, "guided"),
                                (64, 400, "guided"), (128, 1, "guided")]:
        items, sio, nia, of, ob = build_work_items(sub_off, 30, tile, target, split)
        assert np.array_equal(np.sort(of), np.arange(len(items))) and np.array_equal(np.sort(ob), np.arange(nia))
        assert (np.diff(items[of, 2]) <= 0).all() and (np.diff(items[ob, 2]) <= 0).all()
        assert items[:, 2].sum() == sub_off[-1]
        for sp in range(50):
            it = items[sio[sp]:sio[sp + 1]]
            assert (it[:, 0] == sp).all()
            if len(it):
                assert it[0, 1] == sub_off[sp] and it[-1, 1] + it[-1, 2] == sub_off[sp + 1]
                assert (it[:-1, 2] % tile == 0).all()
                assert np.array_equal(it[:, 3], np.arange(len(it)))
        assert nia == sio[30]


def test_launch_records_resolve_the_index_chain():
    """d_launch_* rows = what the kernels would otherwise load through launch order -> work list -> subdomain index."""
    rng = np.random.default_rng(1)
    counts = rng.integers(0, 3000, size=40)
    counts[[3, 17]] = 0                                         # subdomains without pairs have no work item
    sub_off = np.concatenate([[0], np.cumsum(counts)])
    sub_ids = rng.permutation(200)[:40].astype(np.int32)        # position -> global subdomain index
    for target in (1, 600):
        items, sio, nia, of, ob = build_work_items(sub_off, 25, 128, target)
        for order in (of, ob):
            rec = launch_records(items, order, sub_ids)
            assert rec.dtype == np.int32 and rec.shape == (len(order), 4)
            for b, it in enumerate(order):
                sp, first, count, _ = items[it]
                assert tuple(rec[b]) == (first, count, sub_ids[sp], it)
        # one item per active subdomain only when no active subdomain is empty or split
        assert not one_item_per_subdomain(sio, 25, nia)         # subdomains 3 and 17 are empty
    counts = np.full(12, 300)
    sub_off = np.concatenate([[0], np.cumsum(counts)])
    items, sio, nia, of, ob = build_work_items(sub_off, 12, 128, 4)
    assert one_item_per_subdomain(sio, 12, nia)
    items, sio, nia, of, ob = build_work_items(sub_off, 12, 128, 100)    # asks for more items: subdomains are split
    assert nia > 12 and not one_item_per_subdomain(sio, 12, nia)
    # an empty and a split subdomain cancel in the item COUNT, not in the map
    sub_off = np.array([0, 0, 600, 700])
    items, sio, nia, of, ob = build_work_items(sub_off, 3, 128, 3)
    if nia == 3:
        assert not one_item_per_subdomain(sio, 3, nia)


def test_constants_reject_unknown_keys():
    with pytest.raises(KeyError):
        Constants(not_a_key=1)
    c = Constants(seed=3)
    assert c.seed == 3 and c["n_steps"] == 15000


def test_every_baseline_config_has_a_tiled_kernel_instance():
    "the five BASELINE configs (true layer sizes, true jet sets) must run on the tiled family, not the generic fallback"
    from fbpinns_b200.engine import Plan
    import common
    for name, fn in configs.CONFIGS.items():
        c = fn()
        sp, _ = c.problem.init_params(**c.problem_init_kwargs)
        ud, xd = sp["dims"]
        small = common.make_case(configs.CONFIGS[name](**configs.SMALL[name]), seed=0)
        for req in small.required_ujss:
            p = Plan(c.network_init_kwargs["layer_sizes"], JetSpec(req, xd, ud))
            assert p.is_fast, (name, req)
        assert Plan(c.network_init_kwargs["layer_sizes"], JetSpec(((0, ()),), xd, ud)).is_fast
    for ls in ([2, 32, 1], [2, 64, 64, 1]):                       # the sweep networks of config 5
        assert Plan(ls, JetSpec(((0, (0, 0)), (0, (1, 1))), 2, 1)).is_fast
    assert not Plan([2, 24, 1], JetSpec(((0, ()),), 2, 1)).is_fast     # falls back to the generic family
    assert not Plan([2, 32, 32, 32, 1], JetSpec(((0, ()),), 2, 1)).is_fast


def test_tensor_family_selection(monkeypatch):
    """which plans get the tcgen05 family: only H = 32 with two hidden layers and C <= 5; `auto` picks it (forward and reverse) for the
    instance timed on hardware (cfg 5's jets) unless FBP_TC_AUTO=0 (tiled) or =fwd (tensor forward, tiled reverse); asking for it on other plans falls back to auto"""
    from fbpinns_b200.engine import Plan
    poisson = JetSpec(((0, (0, 0)), (0, (1, 1))), 2, 1)
    monkeypatch.delenv("FBP_TC_AUTO", raising=False)
    p = Plan([2, 32, 32, 1], poisson)
    assert p.has_tensor and p.kernel == "auto" and p.forward_family == "tensor" and p.cache_per_pair == 0   # tensor reverse too
    monkeypatch.setenv("FBP_TC_AUTO", "fwd")
    assert Plan([2, 32, 32, 1], poisson).cache_per_pair == 32 * 5        # tensor forward + tiled reverse with the cache
    monkeypatch.delenv("FBP_TC_AUTO")
    p.set_kernel("tiled")
    assert p.forward_family == "tiled"
    p.set_kernel("tensor-full")
    assert p.kernel == "tensor-full" and p.forward_family == "tensor" and p.cache_per_pair == 0      # recomputes, no cache
    monkeypatch.setenv("FBP_TC_AUTO", "0")
    assert Plan([2, 32, 32, 1], poisson).forward_family == "tiled"
    monkeypatch.delenv("FBP_TC_AUTO")
    # other jets of the same network: instance exists, but auto stays tiled until validated
    ho = Plan([1, 32, 32, 1], JetSpec(((0, ()), (0, (0,)), (0, (0, 0))), 1, 1))
    assert ho.has_tensor and ho.forward_family == "tiled"
    ho.set_kernel("tensor")
    assert ho.kernel == "tensor" and ho.forward_family == "tensor"
    for ls, jet in (([2, 64, 64, 1], poisson), ([2, 32, 1], poisson), ([2, 16, 1], poisson),
                    ([3, 32, 32, 1], JetSpec(((0, (0, 0)), (0, (1, 1)), (0, (2, 2))), 3, 1))):       # C = 7 does not fit TMEM
        q = Plan(ls, jet)
        assert not q.has_tensor and q.forward_family == "tiled"
        q.set_kernel("tensor-full")
        assert q.kernel == "auto" and q.forward_family == "tiled"


def test_affine_constraining_fast_path_matches_generic():
    "A(x) u + B(x) detection + Leibniz on static coefficient jets == nested-jvp path (values and reverse mode)"
    from fbpinns_b200.jets import AffineConstraining
    from fbpinns_b200 import problems as P
    torch.manual_seed(0)
    cases = [(P.BurgersEquation2D, 2, ((0, ()), (0, (0,)), (0, (1,)), (0, (0, 0)))),
             (P.WaveEquationGaussianVelocity3D, 3, ((0, (0, 0)), (0, (1, 1)), (0, (2, 2)))),
             (P.HarmonicOscillator1DHardBC, 1, ((0, ()), (0, (0,)), (0, (0, 0)))),
             (P.Poisson2D, 2, ((0, (0, 0)), (0, (1, 1))))]
    for prob, xd, req in cases:
        sp, _ = prob.init_params()
        ap = {"static": {"problem": {k: (v.double() if torch.is_tensor(v) else v) for k, v in sp.items()}}, "trainable": {}}
        x = torch.rand(23, xd, dtype=torch.float64) * 0.8 + 0.1
        jet = JetSpec(req, xd, 1)
        aff = AffineConstraining.build(jet, x, prob.constraining_fn, ap)
        assert aff is not None, prob
        uj = torch.randn(23, jet.C, dtype=torch.float64, requires_grad=True)
        got = aff.ujs(uj)
        ref = jet.ujs_constrained(uj, x, prob.constraining_fn, ap)
        w = [torch.randn_like(r) for r in ref]
        for a, b in zip(got, ref):
            assert torch.allclose(a, b, rtol=1e-9, atol=1e-9)
        g1, = torch.autograd.grad(sum((a * ww).sum() for a, ww in zip(got, w)), uj)
        g2, = torch.autograd.grad(sum((b * ww).sum() for b, ww in zip(ref, w)), uj)
        assert torch.allclose(g1, g2, rtol=1e-9, atol=1e-9)
    # a non-affine operator must be rejected (-> generic path)
    nonaff = lambda ap, x, u: torch.tanh(x[:, 0:1]) * u ** 2
    x = torch.rand(11, 1, dtype=torch.float64)
    assert AffineConstraining.build(JetSpec(((0, ()), (0, (0,))), 1, 1), x, nonaff, {}) is None


def test_umma_descriptors_match_cute(tmp_path):
    """tcgen05 family (csrc/fbp_tc.cuh): the hand-encoded shared-memory / instruction descriptors and the operand
    layout in shared memory are bit-identical with what the CuTe headers of this image define (host-side check)."""
    import glob
    import shutil
    import subprocess
    import site
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = None
    for sp in site.getsitepackages():
        hits = glob.glob(os.path.join(sp, "*", "data", "cutlass", "include")) + glob.glob(os.path.join(sp, "*", "3rdparty", "cutlass", "include"))
        hits = [h for h in hits if os.path.exists(os.path.join(h, "cute", "arch", "mma_sm100_desc.hpp"))]
        if hits:
            inc = hits[0]
            break
    if inc is None or shutil.which("nvcc") is None:
        pytest.skip("no CuTe sm_100 headers / nvcc in this environment")
    exe = str(tmp_path / "umma_check")
    cmd = ["nvcc", "-std=c++17", "-I" + inc, "-I" + os.path.join(root, "fbpinns_b200", "csrc"), "-gencode",
           "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr", "-o", exe, os.path.join(root, "tests", "tools", "umma_desc_check.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_jax_prng_known_answers():
    """util/jax_prng.py (restated jax.random threefry): the three Threefry-2x32 known-answer vectors of the Random123
    reference implementation (the ones JAX's own test-suite checks), two widely published JAX outputs, and the
    consistency of the batched (vmap-like) initialisation with the per-key one."""
    from fbpinns_b200.util import jax_prng as J
    from fbpinns_b200.networks import FCN
    hx = lambda r: tuple(int(v) for v in r)
    assert hx(J.threefry_2x32(np.uint32([0, 0]), np.uint32([0, 0]))) == (0x6b200159, 0x99ba4efe)
    m = 0xffffffff
    assert hx(J.threefry_2x32(np.uint32([m, m]), np.uint32([m, m]))) == (0x1cb996fc, 0xbb002be7)
    assert hx(J.threefry_2x32(np.uint32([0x13198a2e, 0x03707344]), np.uint32([0x243f6a88, 0x85a308d3]))) == (0xc4923a9c, 0x483df7a0)
    assert J.split(J.PRNGKey(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert J.uniform(J.PRNGKey(0)) == np.float32(0.41845703)
    root = J.PRNGKey(3)
    _, batched = FCN.init_params_batched(root, 5, [2, 16, 1])
    subkeys = J.split(root, 6)[1:]
    for i in range(5):
        _, single = FCN.init_params(subkeys[i], [2, 16, 1])
        for (W, B), (w, b) in zip(batched["layers"], single["layers"]):
            assert torch.equal(W[i], w) and torch.equal(B[i], b)
            assert float(w.abs().max()) <= 1 / np.sqrt(w.shape[1]) and w.dtype == torch.float32


def test_checkpoint_is_in_the_reference_pickle_format(tmp_path):
    """util/checkpoint.py: model_{i:08d}.jax = pickle of (i, all_params, (ScaleByAdamState(count, mu, nu), EmptyState()),
    active, u_test_losses) with numpy leaves (fbpinns/trainers_base.py:64-69, trainers.py:721); the optax classes are
    referenced by their real module paths, so the file loads where optax exists and round-trips here without it."""
    import pickle
    import pickletools
    from fbpinns_b200.util import checkpoint as ck
    rng = np.random.default_rng(0)
    layers = [(torch.tensor(rng.standard_normal((3, 4, 2)), dtype=torch.float32), torch.tensor(rng.standard_normal((3, 4)), dtype=torch.float32))]
    all_params = {"static": {"problem": {"dims": (1, 2)}, "decomposition": {"m": 3, "_device_cache": object()}},
                  "trainable": {"network": {"subdomain": {"layers": layers}}, "problem": {"mu": torch.tensor(0.5)}}}
    mu = {"network": {"subdomain": {"layers": [(torch.zeros(3, 4, 2), torch.ones(3, 4))]}}, "problem": {"mu": torch.tensor(0.1)}}
    path = str(tmp_path / "model_00000010.jax")
    ck.save_model(path, 10, all_params, mu, mu, 10, np.array([1, 2, 0]), [[0, 0.1, 0.5, 0.6]])
    raw = open(path, "rb").read()
    names = {arg for op, arg, _ in pickletools.genops(raw) if isinstance(arg, str)}
    assert any("optax._src.transform" in n for n in names) and any("ScaleByAdamState" in n for n in names)
    assert "optax" not in __import__("sys").modules or __import__("sys").modules["optax"].__name__ == "optax"
    i, ap, (count, mu2, nu2), active, losses = ck.load_model(path)
    assert i == 10 and count == 10 and active.tolist() == [1, 2, 0] and losses.shape == (1, 4)
    assert "_device_cache" not in ap["static"]["decomposition"] and ap["static"]["problem"]["dims"] == (1, 2)
    w, b = ap["trainable"]["network"]["subdomain"]["layers"][0]
    assert isinstance(w, np.ndarray) and np.array_equal(w, layers[0][0].numpy()) and isinstance(ap["trainable"]["network"]["subdomain"]["layers"][0], tuple)
    assert np.array_equal(mu2["network"]["subdomain"]["layers"][0][1], np.ones((3, 4), dtype=np.float32))
    # the raw pickle has the reference's outer structure
    with ck.optax_classes():
        model = pickle.loads(raw)
    assert isinstance(model, tuple) and len(model) == 5 and type(model[2][0]).__name__ == "ScaleByAdamState" and model[2][1] == ()


def test_xla_ffi_shim_compiles_against_stub():
    """csrc/xla_ffi_shim.cc (the jax.ffi binding over the C ABI) cannot be built for real here (no jaxlib headers); it is at
    least type-checked against a minimal stand-in of xla/ffi/api/ffi.h, so that its calls stay in step with
    include/fbpinn_b200.h (argument order and count of fbp_forward / fbp_reduce_* / fbp_backward)."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    cuda_inc = "/usr/local/cuda/include"
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I" + os.path.join(ROOT, "tests", "tools", "xla_ffi_stub"),
           "-I" + os.path.join(ROOT, "include"), "-I" + cuda_inc, os.path.join(ROOT, "fbpinns_b200", "csrc", "xla_ffi_shim.cc")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-3000:]
