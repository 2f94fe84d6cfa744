"""
Golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE (/root/reference/fbpinns, read-only, not copied)
under a numpy-backed shim of the small `jax` API subset that source uses.  JAX itself is not installed in this
image, so the reference cannot run natively; the shim supplies
    jax.numpy  -> numpy (with jnp.array's float64->float32 / int64->int32 demotion and the `.at[].set/add` idiom),
    jax.vmap   -> a python loop honouring pytree in_axes,      jax.ops.segment_sum -> np.add.at,
    jax.tree_util.tree_map, jax.nn.tanh/sigmoid, jit -> identity;   jvp / value_and_grad are NOT provided.
What this pins (bit-for-bit the reference's code paths, evaluated in float64 after init):
    * RectangularDecompositionND.init_params / _get_level_params (box arithmetic, float32 casts)
    * get_jmaps
    * get_inputs (active-mask algebra, m_take re-index, unique (point, pou) rows) on the reference's dense
      inside test (`_inside_rectangleND` + nonzero, the specification at decompositions_base.py:100-106)
    * FBPINN_model: norm_fn, FCN.network_fn, unnorm_fn, window_fn (windows.cosine), both segment sums, /wp, /npou,
      constraining_fn -> u, wp, us, ws, us_raw;  derivatives (ujs) by central finite differences of that u
    * Problem.loss_fn / constraining_fn of the reference problems
    * refindex.npz: the reference's own batched inside_points_batch / inside_models_batch (lax.map / lax.scan code,
      checked here against its dense specification for several batch sizes), RectangularDecompositionND.inside_points /
      inside_models, and FBPINNTrainer._get_x_batch + _get_update_inputs (per-constraint split of the takes) for
      several 0/1/2 active masks on a two-constraint setup
The outputs are committed as tests/golden/refmodel_*.npz and checked by tests/test_golden_reference.py against the
oracle (and through it the CUDA path).  Run:  python tests/golden/make_golden_shim.py
"""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_ROOT = "/root/reference"


# ------------------------------------------------------------------------------------------------ the shim

class AtArray(np.ndarray):
    @property
    def at(self):
        return _At(self)


class _At:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        return _AtIdx(self.a, idx)


class _AtIdx:
    def __init__(self, a, idx):
        self.a, self.idx = a, idx

    def set(self, v):
        out = np.array(self.a, copy=True)
        out[self.idx] = v
        return out.view(AtArray)

    def add(self, v):
        out = np.array(self.a, copy=True)
        np.add.at(out, self.idx, v)
        return out.view(AtArray)


def _wrap(x):
    if isinstance(x, np.ndarray):
        return x.view(AtArray)
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _jnp_array(x, dtype=None):
    a = np.array(x, dtype=dtype)
    if dtype is None:                       # JAX defaults with x64 disabled
        if a.dtype == np.float64:
            a = a.astype(np.float32)
        elif a.dtype == np.int64:
            a = a.astype(np.int32)
    return a.view(AtArray)


class _Jnp(types.ModuleType):
    ndarray = np.ndarray
    pi = np.pi

    def __getattr__(self, name):
        a = getattr(np, name)
        if callable(a) and not isinstance(a, type):
            return lambda *args, **kw: _wrap(a(*args, **kw))
        return a


def tree_map(f, tree, *rest, is_leaf=None):
    if is_leaf is not None and is_leaf(tree):
        return f(tree, *rest)
    if tree is None:
        return None
    if isinstance(tree, dict):
        return {k: tree_map(f, tree[k], *[r[k] for r in rest], is_leaf=is_leaf) for k in tree}
    if isinstance(tree, (list, tuple)):
        out = [tree_map(f, t, *[r[i] for r in rest], is_leaf=is_leaf) for i, t in enumerate(tree)]
        return type(tree)(out) if not hasattr(tree, "_fields") else type(tree)(*out)
    return f(tree, *rest)


def _take(a, ax, i):
    if ax is None:
        return a
    if isinstance(ax, int):
        assert ax == 0
        return tree_map(lambda x: x[i], a)
    if isinstance(ax, dict):
        return {k: _take(a[k], ax[k], i) for k in a}
    if isinstance(ax, (list, tuple)):
        return type(a)(_take(x, y, i) for x, y in zip(a, ax))
    raise TypeError(ax)


def _batch_size(a, ax):
    if ax is None:
        return None
    if isinstance(ax, int):
        sizes = []
        tree_map(lambda x: sizes.append(np.shape(x)[0]), a)
        return sizes[0] if sizes else None
    if isinstance(ax, dict):
        for k in a:
            n = _batch_size(a[k], ax[k])
            if n is not None:
                return n
        return None
    for x, y in zip(a, ax):
        n = _batch_size(x, y)
        if n is not None:
            return n
    return None


def vmap(f, in_axes=0):
    def g(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(b for b in (_batch_size(a, ax) for a, ax in zip(args, axes)) if b is not None)
        outs = [f(*[_take(a, ax, i) for a, ax in zip(args, axes)]) for i in range(n)]
        if isinstance(outs[0], tuple):
            return tuple(np.stack([o[j] for o in outs]).view(AtArray) for j in range(len(outs[0])))
        return np.stack(outs).view(AtArray)
    return g


def segment_sum(data, segment_ids, num_segments=None, indices_are_sorted=False):
    out = np.zeros((num_segments,) + tuple(np.shape(data)[1:]), dtype=np.asarray(data).dtype)
    np.add.at(out, np.asarray(segment_ids), np.asarray(data))
    return out.view(AtArray)


def _jit(f=None, **kw):
    if f is None:
        return lambda g: g
    return f


class _Stub(types.ModuleType):
    "module whose every attribute is a harmless dummy (matplotlib, IPython, tensorboardX, optax, PIL ...)"

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        d = _Stub(self.__name__ + "." + name)
        setattr(self, name, d)
        return d

    def __call__(self, *a, **k):
        return _Stub("call")


def install_shim():
    jnp = _Jnp("jax.numpy")
    jnp.array = _jnp_array
    jax = types.ModuleType("jax")
    jax.numpy = jnp
    jax.jit, jax.vmap = _jit, vmap

    def _no(*a, **k):
        raise RuntimeError("autodiff is not provided by the shim")
    jax.jvp = jax.value_and_grad = jax.grad = _no
    jax.nn = types.ModuleType("jax.nn")
    jax.nn.tanh, jax.nn.sigmoid = np.tanh, (lambda x: 1 / (1 + np.exp(-x)))
    jax.tree_util = types.ModuleType("jax.tree_util")
    jax.tree_util.tree_map = tree_map
    jax.tree_util.tree_flatten = _no
    jax.tree_util.tree_leaves = lambda t: (lambda acc: (tree_map(lambda x: acc.append(x), t), acc)[1])([])
    jax.ops = types.ModuleType("jax.ops")
    jax.ops.segment_sum = segment_sum
    jax.random = _Stub("jax.random")
    jax.lax = types.ModuleType("jax.lax")

    def lax_map(f, xs):
        n = len(xs[0]) if isinstance(xs, (tuple, list)) else len(xs)
        outs = [f(tuple(_wrap(np.asarray(x[i])) for x in xs) if isinstance(xs, (tuple, list)) else _wrap(np.asarray(xs[i])))
                for i in range(n)]
        if isinstance(outs[0], tuple):
            return tuple(np.stack([np.asarray(o[j]) for o in outs]).view(AtArray) for j in range(len(outs[0])))
        return np.stack([np.asarray(o) for o in outs]).view(AtArray)

    def lax_scan(f, init, xs):
        carry = init
        n = len(xs[0]) if isinstance(xs, (tuple, list)) else len(xs)
        ys = []
        for i in range(n):
            xi = tuple(_wrap(np.asarray(x[i])) for x in xs) if isinstance(xs, (tuple, list)) else _wrap(np.asarray(xs[i]))
            carry, y = f(carry, xi)
            ys.append(y)
        return carry, (None if ys and ys[0] is None else ys)

    def lax_dynamic_slice(x, start, sizes):
        idx = tuple(slice(int(s0), int(s0) + int(sz)) for s0, sz in zip(start, sizes))
        return _wrap(np.asarray(x)[idx])
    jax.lax.map, jax.lax.scan, jax.lax.dynamic_slice = lax_map, lax_scan, lax_dynamic_slice
    mods = {"jax": jax, "jax.numpy": jnp, "jax.nn": jax.nn, "jax.tree_util": jax.tree_util, "jax.ops": jax.ops,
            "jax.random": jax.random, "jax.lax": jax.lax}
    for name in ["optax", "matplotlib", "matplotlib.pyplot", "matplotlib.collections", "IPython", "IPython.display",
                 "tensorboardX"]:
        mods[name] = _Stub(name)
    sys.modules.update(mods)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ref = types.SimpleNamespace()
    ref.trainers = importlib.import_module("fbpinns.trainers")
    ref.decompositions = importlib.import_module("fbpinns.decompositions")
    ref.networks = importlib.import_module("fbpinns.networks")
    ref.problems = importlib.import_module("fbpinns.problems")
    ref.windows = importlib.import_module("fbpinns.windows")
    ref.trainers.logger.setLevel("ERROR")
    return ref


# ------------------------------------------------------------------------------------------------ cases

def f64(tree):
    return tree_map(lambda x: np.asarray(x, dtype=np.float64).view(AtArray) if isinstance(x, np.ndarray) and x.dtype.kind == "f" else x, tree)


class DenseDecomposition:
    "the reference's decomposition with inside_points evaluated as its self-test specifies (dense mask + nonzero)"

    def __init__(self, ref_cls):
        self.ref_cls = ref_cls

    def inside_points(self, all_params, x_batch):
        ps = {"params": all_params["static"]["decomposition"]["subdomain"]["params"]}
        m = all_params["static"]["decomposition"]["m"]
        ps32 = {"params": [np.asarray(p, dtype=np.float32) for p in ps["params"]]}
        inside = self.ref_cls._inside_rectangleND(ps32, np.asarray(x_batch, dtype=np.float32), np.arange(m))
        n_take, m_take = np.nonzero(inside)
        return _wrap(n_take), _wrap(m_take), _wrap(np.nonzero(np.any(inside, axis=0))[0])


def make_case(ref, name, rng):
    sys.path.insert(0, HERE)
    from cases import case_setup
    cs = case_setup(name)
    dkw, x, layer_sizes, req = cs["dkw"], cs["x"], cs["layer_sizes"], cs["req"]
    prob, pkw = getattr(ref.problems, cs["problem"]), cs["pkw"]
    dstat, _ = ref.decompositions.RectangularDecompositionND.init_params(**dkw)
    pstat, ptrain = prob.init_params(**pkw)
    m = dstat["m"]
    layers = []
    for fi, fo in zip(layer_sizes[:-1], layer_sizes[1:]):
        v = np.sqrt(1 / fi)
        layers.append((rng.uniform(-v, v, (m, fo, fi)).astype(np.float32), rng.uniform(-v, v, (m, fo)).astype(np.float32)))
    return dict(name=name, dkw=dkw, prob=prob, pstat=pstat, ptrain=ptrain, dstat=dstat, x=x, layers=layers,
                layer_sizes=layer_sizes, req=req, m=m)


def run_case(ref, case, rng):
    T = ref.trainers
    dcls = ref.decompositions.RectangularDecompositionND
    m, x32 = case["m"], case["x"]
    out = {}
    # --- A1: level params (float64) and the float32 static leaves
    lp = dcls._get_level_params(0, len(case["dkw"]["subdomain_xs"]), case["dkw"]["subdomain_xs"], case["dkw"]["subdomain_ws"],
                                case["dkw"]["unnorm"])
    for i, p in enumerate(lp):
        out[f"level_{i}"] = np.asarray(p)
    for i, p in enumerate(case["dstat"]["subdomain"]["params"]):
        out[f"static_{i}"] = np.asarray(p)
    out["xmins0"], out["xmaxs0"] = np.asarray(case["dstat"]["xmins0"]), np.asarray(case["dstat"]["xmaxs0"])
    # --- jmaps
    nodes, leaves, jac_is = T.get_jmaps(case["req"])
    out["jmaps_repr"] = np.array(repr((nodes, leaves, jac_is)))
    # --- A3: get_inputs with a mixed active mask
    active = rng.integers(0, 3, size=m)
    active[:2] = 1
    all_params32 = {"static": {"decomposition": case["dstat"], "problem": case["pstat"]}, "trainable": {}}
    takes, all_ims, (active2, cut_active, cut_fixed, cut_all, merge_active) = T.get_inputs(_wrap(x32), active, all_params32,
                                                                                          DenseDecomposition(dcls))
    out["active_in"], out["active_out"], out["all_ims"] = active, np.asarray(active2), np.asarray(all_ims)
    for nm, a in zip(["m_take", "n_take", "p_take", "np_take"], takes[:4]):
        out[nm] = np.asarray(a)
    out["npou"] = np.array(takes[4])
    # --- A5/A6: FBPINN_model in float64 (float32-rounded static leaves and parameters, upcast)
    layers64 = [(np.asarray(w, np.float64).view(AtArray), np.asarray(b, np.float64).view(AtArray)) for w, b in case["layers"]]
    full = {"static": {"decomposition": f64(case["dstat"]), "problem": f64(case["pstat"])},
            "trainable": {"network": {"subdomain": {"layers": layers64}}}}
    if case["ptrain"]:
        full["trainable"]["problem"] = f64(case["ptrain"])
    cut = {"static": cut_all(full["static"]), "trainable": cut_all(full["trainable"])}
    model_fns = (dcls.norm_fn, ref.networks.FCN.network_fn, dcls.unnorm_fn, dcls.window_fn, case["prob"].constraining_fn)
    x64 = np.asarray(x32, np.float64).view(AtArray)

    def model(xb):
        return T.FBPINN_model(cut, xb, takes, model_fns, verbose=False)
    u, wp, us, ws, us_raw = model(x64)
    out.update(u=np.asarray(u), wp=np.asarray(wp), us=np.asarray(us), ws=np.asarray(ws), us_raw=np.asarray(us_raw))
    # unconstrained twin (identity constraining operator)
    fns_id = model_fns[:4] + (ref.problems.Problem.constraining_fn,)
    out["u_unconstrained"] = np.asarray(T.FBPINN_model(cut, x64, takes, fns_id, verbose=False)[0])
    # --- A7: ujs by central finite differences of the reference's u (points move, takes stay fixed)
    xd = x64.shape[1]
    ps = case["dstat"]["subdomain"]["params"]
    sd_min = float(np.min((np.asarray(ps[1], np.float64) - np.asarray(ps[0], np.float64)) / 2))
    h = 1e-4 * sd_min
    out["fd_h"] = np.array(h)

    def fd(path):
        if len(path) == 0:
            return np.asarray(u)
        e = np.zeros(xd)
        e[path[0]] = h
        if len(path) == 1:
            return (np.asarray(model((x64 + e).view(AtArray))[0]) - np.asarray(model((x64 - e).view(AtArray))[0])) / (2 * h)
        assert path[0] == path[1]
        return (np.asarray(model((x64 + e).view(AtArray))[0]) - 2 * np.asarray(u) + np.asarray(model((x64 - e).view(AtArray))[0])) / h ** 2
    for j, (iu, path) in enumerate(case["req"]):
        out[f"uj_fd_{j}"] = fd(path)[:, iu:iu + 1]
    # --- loss_fn of the reference problem on the FD ujs (pins the torch restatement of loss_fn)
    cons = [[x64] + [out[f"uj_fd_{j}"].view(AtArray) for j in range(len(case["req"]))]]
    try:
        out["loss_on_fd_ujs"] = np.array(float(case["prob"].loss_fn(cut, cons)))
    except Exception as e:                     # problems with more than one constraint are exercised elsewhere
        out["loss_on_fd_ujs"] = np.array(np.nan)
    # --- reverse mode: central finite differences over PARAMETERS of L0 = sum_p R_p u(x_p), u from the reference's
    #     FBPINN_model (constrained), for a handful of parameter entries of every layer
    R = rng.normal(size=np.asarray(u).shape)
    out["grad_R"] = R
    picks, fds = [], []
    for l, (w, b) in enumerate(layers64):
        for which, arr in (("w", w), ("b", b)):
            for _ in range(3):
                idx = tuple(int(rng.integers(0, d)) for d in arr.shape)
                old = float(arr[idx])
                hp = 1e-6 * max(1.0, abs(old))
                vals = []
                for sgn in (+1, -1):
                    arr[idx] = old + sgn * hp
                    cut_p = {"static": cut["static"], "trainable": cut_all(full["trainable"])}
                    vals.append(float((R * np.asarray(T.FBPINN_model(cut_p, x64, takes, model_fns, verbose=False)[0])).sum()))
                arr[idx] = old
                picks.append((l, 0 if which == "w" else 1) + idx + (0,) * (3 - len(idx)))
                fds.append((vals[0] - vals[1]) / (2 * hp))
    out["grad_picks"], out["grad_fd"] = np.array(picks), np.array(fds)
    # inputs
    out["x"] = x32
    for l, (w, b) in enumerate(case["layers"]):
        out[f"W{l}"], out[f"b{l}"] = w, b
    out["layer_sizes"] = np.array(case["layer_sizes"])
    out["req_repr"] = np.array(repr(case["req"]))
    return out


def run_index_case(ref, rng):
    """A2-A4 with the reference's OWN batched inside tests (decompositions_base.py lax.map / lax.scan code) and its
    own FBPINNTrainer._get_update_inputs (per-constraint split of the takes), on HarmonicOscillator1D (two
    constraints) and BurgersEquation2D with scheduler-like 0/1/2 active masks."""
    T = ref.trainers
    dcls = ref.decompositions.RectangularDecompositionND
    base = importlib.import_module("fbpinns.decompositions_base")
    sys.path.insert(0, HERE)
    from cases import case_setup
    out = {}
    for tag, name in [("ho", "ho1d_hardbc"), ("bg", "burgers2d")]:
        cs = case_setup(name)
        dstat, _ = dcls.init_params(**cs["dkw"])
        m = dstat["m"]
        x = cs["x"]
        all_params = {"static": {"decomposition": dstat, "problem": {"dims": (1, x.shape[1])}}, "trainable": {}}
        # the reference's batched implementation for several batch sizes against its own dense specification
        ps = {"params": dstat["subdomain"]["params"]}
        dense = np.asarray(dcls._inside_rectangleND(ps, _wrap(x), np.arange(m)))
        nt, mt = np.nonzero(dense)
        for bs in [1, 7, len(x)]:
            n_take, m_take, inside_ims = base.inside_points_batch(ps, _wrap(x), _wrap(np.arange(m)), bs, dcls._inside_rectangleND)
            assert np.array_equal(np.asarray(n_take), nt) and np.array_equal(np.asarray(m_take), mt), (tag, bs)
            sel = _wrap(np.arange(0, m, 2))
            ips, d = base.inside_models_batch(ps, _wrap(x), sel, bs, dcls._inside_rectangleND)
            assert np.array_equal(np.asarray(ips), np.nonzero(dense[:, ::2].any(1))[0]), (tag, bs)
        n_take, m_take, inside_ims = dcls.inside_points(all_params, _wrap(x))
        out[f"{tag}_x"], out[f"{tag}_n_take"], out[f"{tag}_m_take"] = x, np.asarray(n_take), np.asarray(m_take)
        out[f"{tag}_inside_ims"] = np.asarray(inside_ims)
        ips, d = dcls.inside_models(all_params, _wrap(x), _wrap(np.arange(1, m, 3)))
        out[f"{tag}_models_sel"], out[f"{tag}_inside_ips"], out[f"{tag}_d"] = np.arange(1, m, 3), np.asarray(ips), np.array(float(d))

    # _get_update_inputs on a two-constraint problem (physics grid + extra points), several active masks
    cs = case_setup("burgers2d")
    dstat, _ = dcls.init_params(**cs["dkw"])
    m = dstat["m"]
    x1 = cs["x"]
    x2 = rng.uniform([-1, 0], [1, 1], size=(17, 2)).astype(np.float32)
    v2 = rng.normal(size=(17, 1)).astype(np.float32)
    constraints_global = [[_wrap(x1)], [_wrap(x2), _wrap(v2)]]
    x_batch_global = _wrap(np.concatenate([x1, x2]))
    offsets = _wrap(np.array([0, len(x1)]))
    fs = np.zeros((len(x1) + len(x2), 2), dtype=bool)
    fs[:len(x1), 0] = True
    fs[len(x1):, 1] = True
    trainable = {"network": {"subdomain": {"layers": [(_wrap(np.zeros((m, 2, 2), np.float32)), _wrap(np.zeros((m, 2), np.float32)))]}}}
    all_params = {"static": {"decomposition": dstat, "problem": {"dims": (1, 2)}}, "trainable": trainable}
    dummy = types.SimpleNamespace(c=types.SimpleNamespace(n_steps=1))
    dummy._get_x_batch = lambda *a: T.FBPINNTrainer._get_x_batch(dummy, *a)
    out["ui_x1"], out["ui_x2"], out["ui_v2"] = x1, x2, v2
    masks = []
    for trial in range(4):
        active = rng.integers(0, 3, size=m)
        active[rng.integers(0, m)] = 1
        if trial == 0:
            active[:] = 1
        masks.append(active.copy())
        (active2, merge_active, active_opt_states, active_params, fixed_params, static_params, takess, constraints, x_batch) = \
            T.FBPINNTrainer._get_update_inputs(dummy, 0, active, all_params, trainable, x_batch_global, constraints_global,
                                               _wrap(fs), offsets, dcls, None)
        out[f"ui{trial}_active_in"], out[f"ui{trial}_active_out"] = active, np.asarray(active2)
        out[f"ui{trial}_x_batch"] = np.asarray(x_batch)
        out[f"ui{trial}_n_active_params"] = np.array(np.asarray(active_params["network"]["subdomain"]["layers"][0][0]).shape[0])
        out[f"ui{trial}_n_fixed_params"] = np.array(np.asarray(fixed_params["network"]["subdomain"]["layers"][0][0]).shape[0])
        for ic, tk in enumerate(takess):
            for nm, a in zip(["m_take", "n_take", "p_take", "np_take"], tk[:4]):
                out[f"ui{trial}_c{ic}_{nm}"] = np.asarray(a)
            out[f"ui{trial}_c{ic}_npou"] = np.array(tk[4])
            for j, c_ in enumerate(constraints[ic]):
                out[f"ui{trial}_c{ic}_arr{j}"] = np.asarray(c_)
    out["ui_trials"] = np.array(len(masks))
    np.savez_compressed(os.path.join(HERE, "refindex.npz"), **out)
    print("refindex.npz:", len(out), "arrays;", "pairs per trial:", [len(out[f"ui{t}_c0_m_take"]) for t in range(len(masks))])


def main():
    ref = install_shim()
    run_index_case(ref, np.random.default_rng(7))
    for i, name in enumerate(["ho1d_hardbc", "burgers2d", "wave3d"]):
        rng = np.random.default_rng(100 + i)
        case = make_case(ref, name, rng)
        out = run_case(ref, case, rng)
        np.savez_compressed(os.path.join(HERE, f"refmodel_{name}.npz"), **out)
        print(name, "pairs", len(out["m_take"]), "points", len(out["x"]), "u range", float(out["u"].min()), float(out["u"].max()))


if __name__ == "__main__":
    main()
