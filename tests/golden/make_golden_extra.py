"""
Second batch of golden vectors from the REFERENCE'S OWN SOURCE run under the numpy jax-shim of make_golden_shim.py
(VERDICT r1 item 6).  Pins what the first batch left oracle-vs-CUDA only:

  * refextra_multilevel.npz — MultilevelRectangularDecompositionND.init_params (fbpinns/decompositions.py:338-375:
    two levels, level 0 a single window-less subdomain, pou ids), get_inputs on it with a mixed 0/1/2 mask
    (npou = 2: the (point, pou) unique rows of trainers.py:380-388), FBPINN_model with the per-level quotient and
    the /npou average (:163-170), finite-difference ujs and parameter finite differences of a value functional;
  * refextra_ho1d.npz — HarmonicOscillator1D's two-constraint soft-BC loss (problems.py:116-132) incl. its empty
    boundary branch, HarmonicOscillator1DInverse (problems.py:209-271): init_params, exact_solution / the 13 data
    points, loss_fn with a trainable mu, dL/dmu by central differences; both on ujs taken as finite differences of the
    reference's own FBPINN_model over the constraint points (the takes come from the reference's _get_update_inputs,
    so the per-constraint split of a physics + boundary/data problem is pinned too).

Run:  python tests/golden/make_golden_extra.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_shim as S                                   # noqa: E402
from make_golden_shim import _wrap, AtArray, f64                # noqa: E402
import cases_extra as CX                                        # noqa: E402


def _fd_ujs(model, x64, req, h):
    "central differences of u = model(x)[0] over the points (takes fixed), one (n,1) array per required (iu, path)"
    u0 = np.asarray(model(x64))
    out = []
    for iu, path in req:
        if len(path) == 0:
            v = u0
        else:
            e = np.zeros(x64.shape[1])
            e[path[0]] = h
            up, um = np.asarray(model((x64 + e).view(AtArray))), np.asarray(model((x64 - e).view(AtArray)))
            if len(path) == 1:
                v = (up - um) / (2 * h)
            else:
                assert len(path) == 2 and path[0] == path[1]
                v = (up - 2 * u0 + um) / h ** 2
        out.append(v[:, iu:iu + 1])
    return out


def run_multilevel(ref, rng):
    T = ref.trainers
    dcls = ref.decompositions.MultilevelRectangularDecompositionND
    cs = CX.multilevel_setup()
    dstat, _ = dcls.init_params(**cs["dkw"])
    m, x32, req = dstat["m"], cs["x"], cs["req"]
    out = {"m": np.array(m), "x": x32, "req_repr": np.array(repr(req)), "layer_sizes": np.array(cs["layer_sizes"])}
    for i, p in enumerate(dstat["subdomain"]["params"]):
        out[f"static_{i}"] = np.asarray(p)
    out["pou"] = np.asarray(dstat["subdomain"]["pou"])
    out["xmins0"], out["xmaxs0"] = np.asarray(dstat["xmins0"]), np.asarray(dstat["xmaxs0"])
    layers = []
    for fi, fo in zip(cs["layer_sizes"][:-1], cs["layer_sizes"][1:]):
        v = np.sqrt(1 / fi)
        layers.append((rng.uniform(-v, v, (m, fo, fi)).astype(np.float32), rng.uniform(-v, v, (m, fo)).astype(np.float32)))
    for l, (w, b) in enumerate(layers):
        out[f"W{l}"], out[f"b{l}"] = w, b
    for trial, active in enumerate(CX.multilevel_masks(m)):
        all_params32 = {"static": {"decomposition": dstat, "problem": {"dims": (1, 2)}}, "trainable": {}}
        takes, all_ims, (active2, cut_active, cut_fixed, cut_all, merge_active) = T.get_inputs(
            _wrap(x32), active, all_params32, S.DenseDecomposition(dcls))
        t = f"t{trial}_"
        out[t + "active_in"], out[t + "active_out"], out[t + "all_ims"] = active, np.asarray(active2), np.asarray(all_ims)
        for nm, a in zip(["m_take", "n_take", "p_take", "np_take"], takes[:4]):
            out[t + nm] = np.asarray(a)
        out[t + "npou"] = np.array(takes[4])
        layers64 = [(np.asarray(w, np.float64).view(AtArray), np.asarray(b, np.float64).view(AtArray)) for w, b in layers]
        full = {"static": {"decomposition": f64(dstat), "problem": {"dims": (1, 2)}},
                "trainable": {"network": {"subdomain": {"layers": layers64}}}}
        cut = {"static": cut_all(full["static"]), "trainable": cut_all(full["trainable"])}
        fns = (dcls.norm_fn, ref.networks.FCN.network_fn, dcls.unnorm_fn, dcls.window_fn, ref.problems.Problem.constraining_fn)
        x64 = np.asarray(x32, np.float64).view(AtArray)
        u, wp, us, ws, us_raw = T.FBPINN_model(cut, x64, takes, fns, verbose=False)
        out.update({t + "u": np.asarray(u), t + "wp": np.asarray(wp), t + "us": np.asarray(us), t + "ws": np.asarray(ws),
                    t + "us_raw": np.asarray(us_raw)})
        ps = dstat["subdomain"]["params"]
        h = 1e-4 * float(np.min((np.asarray(ps[1], np.float64) - np.asarray(ps[0], np.float64)) / 2))
        out[t + "fd_h"] = np.array(h)
        for j, v in enumerate(_fd_ujs(lambda xb: T.FBPINN_model(cut, xb, takes, fns, verbose=False)[0], x64, req, h)):
            out[t + f"uj_fd_{j}"] = v
        # parameter finite differences of L0 = sum_p R_p u(x_p)
        R = rng.normal(size=np.asarray(u).shape)
        out[t + "grad_R"] = R
        picks, fds = [], []
        for l, (w, b) in enumerate(layers64):
            for which, arr in (("w", w), ("b", b)):
                for _ in range(4):
                    idx = tuple(int(rng.integers(0, d)) for d in arr.shape)
                    old = float(arr[idx])
                    hp = 1e-6 * max(1.0, abs(old))
                    vals = []
                    for sgn in (+1, -1):
                        arr[idx] = old + sgn * hp
                        cut_p = {"static": cut["static"], "trainable": cut_all(full["trainable"])}
                        vals.append(float((R * np.asarray(T.FBPINN_model(cut_p, x64, takes, fns, verbose=False)[0])).sum()))
                    arr[idx] = old
                    picks.append((l, 0 if which == "w" else 1) + idx + (0,) * (3 - len(idx)))
                    fds.append((vals[0] - vals[1]) / (2 * hp))
        out[t + "grad_picks"], out[t + "grad_fd"] = np.array(picks), np.array(fds)
    out["n_trials"] = np.array(trial + 1)
    np.savez_compressed(os.path.join(HERE, "refextra_multilevel.npz"), **out)
    print("refextra_multilevel.npz: m", m, "trials", trial + 1, "pairs", [len(out[f"t{i}_m_take"]) for i in range(trial + 1)],
          "rows", [len(out[f"t{i}_np_take"]) for i in range(trial + 1)])


def run_ho1d(ref, rng):
    T = ref.trainers
    P = ref.problems
    dcls = ref.decompositions.RectangularDecompositionND
    cs = CX.ho1d_setup()
    dstat, _ = dcls.init_params(**cs["dkw"])
    m = dstat["m"]
    out = {"m": np.array(m), "layer_sizes": np.array(cs["layer_sizes"])}
    layers = []
    for fi, fo in zip(cs["layer_sizes"][:-1], cs["layer_sizes"][1:]):
        v = np.sqrt(1 / fi)
        layers.append((rng.uniform(-v, v, (m, fo, fi)).astype(np.float32), rng.uniform(-v, v, (m, fo)).astype(np.float32)))
    for l, (w, b) in enumerate(layers):
        out[f"W{l}"], out[f"b{l}"] = w, b
    layers64 = [(np.asarray(w, np.float64).view(AtArray), np.asarray(b, np.float64).view(AtArray)) for w, b in layers]
    domain = types.SimpleNamespace(sample_interior=lambda all_params, key, sampler, batch_shape: _wrap(cs["x_phys"]))
    fns = lambda prob: (dcls.norm_fn, ref.networks.FCN.network_fn, dcls.unnorm_fn, dcls.window_fn, prob.constraining_fn)
    dummy = types.SimpleNamespace(c=types.SimpleNamespace(n_steps=1))
    dummy._get_x_batch = lambda *a: T.FBPINNTrainer._get_x_batch(dummy, *a)

    for tag, prob, pkw in [("soft", P.HarmonicOscillator1D, cs["pkw"]), ("inv", P.HarmonicOscillator1DInverse, cs["pkw"])]:
        pstat, ptrain = prob.init_params(**pkw)
        for k_, v_ in pstat.items():
            out[f"{tag}_pstat_{k_}"] = np.asarray(v_)
        for k_, v_ in (ptrain or {}).items():
            out[f"{tag}_ptrain_{k_}"] = np.asarray(v_)
        ap0 = {"static": {"decomposition": dstat, "problem": pstat}, "trainable": {"problem": ptrain} if ptrain else {}}
        cons = prob.sample_constraints(ap0, domain, None, "grid", ((len(cs["x_phys"]),),))
        reqs = [c_[-1] for c_ in cons]
        out[f"{tag}_reqs_repr"] = np.array(repr(reqs))
        cons_arr = [[np.asarray(a) for a in c_[:-1]] for c_ in cons]
        for ic, c_ in enumerate(cons_arr):
            for j, a in enumerate(c_):
                out[f"{tag}_c{ic}_arr{j}"] = a
        x_batch_global = _wrap(np.concatenate([c_[0] for c_ in cons_arr]).astype(np.float32))
        sizes = [len(c_[0]) for c_ in cons_arr]
        offsets = _wrap(np.array([0] + list(np.cumsum(sizes)[:-1])))
        fs = np.zeros((sum(sizes), len(sizes)), dtype=bool)
        o = 0
        for ic, s_ in enumerate(sizes):
            fs[o:o + s_, ic] = True
            o += s_
        trainable = {"network": {"subdomain": {"layers": [(_wrap(w), _wrap(b)) for w, b in layers]}}}
        if ptrain:
            trainable["problem"] = ptrain
        all_params = {"static": {"decomposition": dstat, "problem": pstat}, "trainable": trainable}
        for trial, active in enumerate(CX.ho1d_masks(m)):
            t = f"{tag}{trial}_"
            (active2, merge_active, active_opt_states, active_params, fixed_params, static_params, takess, constraints,
             x_batch) = T.FBPINNTrainer._get_update_inputs(
                dummy, 0, active, all_params, trainable, x_batch_global,
                [[_wrap(np.asarray(a, np.float32)) for a in c_] for c_ in cons_arr], _wrap(fs), offsets, dcls, None)
            out[t + "active_in"], out[t + "active_out"] = active, np.asarray(active2)
            # evaluate the reference model per constraint in float64 on the cut trees the trainer would concatenate
            mu_val = 0.7
            stat64 = f64(static_params)
            act64, fix64 = f64(active_params), f64(fixed_params)
            cat = lambda a, b: np.concatenate([np.asarray(a), np.asarray(b)], 0).view(AtArray)
            lay = [(cat(aw, fw), cat(ab, fb)) for (aw, ab), (fw, fb) in
                   zip(act64["network"]["subdomain"]["layers"], fix64["network"]["subdomain"]["layers"])]
            ap = {"static": stat64, "trainable": {"network": {"subdomain": {"layers": lay}}}}
            if ptrain:
                ap["trainable"]["problem"] = {"mu": np.float64(mu_val)}
            cons_eval = []
            for ic, (tk, con) in enumerate(zip(takess, constraints)):
                for nm, a in zip(["m_take", "n_take", "p_take", "np_take"], tk[:4]):
                    out[t + f"c{ic}_{nm}"] = np.asarray(a)
                x64 = np.asarray(con[0], np.float64).view(AtArray)
                out[t + f"c{ic}_x"] = np.asarray(con[0])
                for j, a in enumerate(con[1:]):
                    out[t + f"c{ic}_arr{j + 1}"] = np.asarray(a)
                if len(x64) == 0:
                    ujs = [np.zeros((0, 1)) for _ in reqs[ic]]
                else:
                    ujs = _fd_ujs(lambda xb: T.FBPINN_model(ap, xb, tk, fns(prob), verbose=False)[0], x64, reqs[ic], 1e-5)
                for j, v in enumerate(ujs):
                    out[t + f"c{ic}_uj_fd_{j}"] = v
                cons_eval.append([x64] + [np.asarray(a, np.float64).view(AtArray) for a in con[1:]] + [v.view(AtArray) for v in ujs])
            out[t + "loss"] = np.array(float(prob.loss_fn(ap, cons_eval)))
            if ptrain:
                hm = 1e-4
                lp = float(prob.loss_fn({**ap, "trainable": {**ap["trainable"], "problem": {"mu": np.float64(mu_val + hm)}}}, cons_eval))
                lm = float(prob.loss_fn({**ap, "trainable": {**ap["trainable"], "problem": {"mu": np.float64(mu_val - hm)}}}, cons_eval))
                out[t + "dloss_dmu_fd"], out[t + "mu"] = np.array((lp - lm) / (2 * hm)), np.array(mu_val)
        out[f"{tag}_n_trials"] = np.array(trial + 1)
        # exact solution of the reference on a few points (pins the torch restatement used for the data constraint)
        xe = np.linspace(0, 1, 7).reshape(-1, 1)
        out[f"{tag}_exact_x"], out[f"{tag}_exact_u"] = xe, np.asarray(prob.exact_solution(ap0, _wrap(xe)))
    np.savez_compressed(os.path.join(HERE, "refextra_ho1d.npz"), **out)
    print("refextra_ho1d.npz:", len(out), "arrays; losses",
          [float(out[f"{tg}{i}_loss"]) for tg in ("soft", "inv") for i in range(int(out[f"{tg}_n_trials"]))])


def main():
    ref = S.install_shim()
    run_multilevel(ref, np.random.default_rng(21))
    run_ho1d(ref, np.random.default_rng(22))


if __name__ == "__main__":
    main()
