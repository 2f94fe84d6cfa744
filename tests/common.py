"""Shared case builder for the parity tests: one Constants -> explicit inputs for both the oracle (CPU) and the
CUDA path (same seeded parameters, same float32 points)."""
import numpy as np
import torch

from oracle import ref_takes, ref_model, ref_step
from fbpinns_b200.jets import JetSpec


class Case:
    pass


def _np(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def make_case(c, seed=0, active=None, multilevel=False):
    """Host-only setup (no GPU needed)."""
    k = Case()
    k.c = c
    rng = np.random.default_rng(seed)
    sd, _ = c.decomposition.init_params(**c.decomposition_init_kwargs)
    sdom, _ = c.domain.init_params(**c.domain_init_kwargs)
    sp, tp = c.problem.init_params(**c.problem_init_kwargs)
    k.all_params = {"static": {"domain": sdom, "problem": sp, "decomposition": sd}, "trainable": {}}
    if tp:
        k.all_params["trainable"]["problem"] = tp
    k.m, k.xd = sd["m"], sd["xd"]
    k.ud = sp["dims"][0]
    k.layer_sizes = list(c.network_init_kwargs["layer_sizes"])
    cons = c.problem.sample_constraints(all_params=k.all_params, domain=c.domain, key=np.random.default_rng(seed + 1),
                                        sampler=c.sampler, batch_shapes=c.ns)
    k.required_ujss = [con[-1] for con in cons]
    k.constraints_global = [[_np(t).astype(np.float32) for t in con[:-1]] for con in cons]
    k.x_batch_global = np.concatenate([con[0] for con in k.constraints_global])
    sizes = [con[0].shape[0] for con in k.constraints_global]
    k.offsets, k.fs = ref_takes.constraint_tables(sizes)
    k.jets = [JetSpec(r, k.xd, k.ud) for r in k.required_ujss]
    k.jmapss = [ref_model.get_jmaps(r) for r in k.required_ujss]
    k.layers = ref_model.init_fcn_params(rng, k.m, k.layer_sizes)           # numpy float32, (m, out, in)
    k.prob_trainable = {kk: _np(v).astype(np.float32) for kk, v in (tp or {}).items()}
    if k.prob_trainable:                                                  # make the inverse problem non-trivial
        k.prob_trainable = {kk: (v + np.float32(0.7)) for kk, v in k.prob_trainable.items()}
    # oracle decomposition (independent restatement)
    if multilevel:
        k.decomp_np = ref_takes.multilevel_init_params(**c.decomposition_init_kwargs)
    else:
        k.decomp_np = ref_takes.rectangular_init_params(**c.decomposition_init_kwargs)
    k.active = np.ones(k.m, dtype=int) if active is None else np.asarray(active)
    k.ui = ref_takes.get_update_inputs(k.active, k.decomp_np, k.x_batch_global, k.constraints_global, k.fs, k.offsets)
    return k


def oracle_all_params(k, dtype):
    """all_params builder for loss_fn / constraining_fn on the oracle side (torch CPU, given dtype)."""
    def conv(v):
        return v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v
    static = {tag: {kk: conv(v) for kk, v in d.items()} for tag, d in k.all_params["static"].items()
              if tag != "decomposition"}

    def make(layers_cut, pt):
        return {"static": static, "trainable": {"problem": pt}}
    return make


def oracle_loss_and_grads(k, dtype=torch.float64, layers=None, prob_trainable=None):
    """FBPINN_loss value + grads w.r.t. the active subdomains' layers and the problem trainables."""
    layers = k.layers if layers is None else layers
    pt = k.prob_trainable if prob_trainable is None else prob_trainable
    ui = k.ui
    decomp_t = ref_model.to_torch(k.decomp_np, dtype)
    decomp_cut = ref_model.cut_decomp(decomp_t, ui["all_ims"])
    al = [(w[ui["active_ims"]], b[ui["active_ims"]]) for w, b in layers]
    fl = [(w[ui["fixed_ims"]], b[ui["fixed_ims"]]) for w, b in layers]
    from fbpinns_b200.problems import Problem
    cf = k.c.problem.constraining_fn if k.c.problem.constraining_fn is not Problem.constraining_fn else None
    return ref_step.loss_and_grads(al, fl, pt, decomp_cut, ui["takess"], ui["constraints"], k.jmapss,
                                   k.c.problem.loss_fn, cf, oracle_all_params(k, dtype), dtype)


def oracle_ujs(k, ic, dtype=torch.float64, constrained=True, layers=None):
    """ujs of constraint ic from the oracle (list of (n,1) numpy arrays)."""
    layers = k.layers if layers is None else layers
    ui = k.ui
    decomp_cut = ref_model.cut_decomp(ref_model.to_torch(k.decomp_np, dtype), ui["all_ims"])
    lc = [(torch.as_tensor(w[ui["all_ims"]], dtype=dtype), torch.as_tensor(b[ui["all_ims"]], dtype=dtype)) for w, b in layers]
    x = torch.as_tensor(ui["constraints"][ic][0], dtype=dtype)
    pt = {kk: torch.as_tensor(v, dtype=dtype) for kk, v in k.prob_trainable.items()}
    ap = oracle_all_params(k, dtype)(lc, pt)
    from fbpinns_b200.problems import Problem
    cf = k.c.problem.constraining_fn if (constrained and k.c.problem.constraining_fn is not Problem.constraining_fn) else None
    ujs = ref_model.fbpinn_forward(decomp_cut, lc, x, ui["takess"][ic], k.jmapss[ic], cf, ap)
    return [u.detach().numpy() for u in ujs]


def rel_err(a, b):
    """max |a-b| / max |b|  (the north-star tolerance: relative to each quantity's max magnitude)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.max(np.abs(b)) if b.size else 1.0
    return float(np.max(np.abs(a - b)) / max(scale, 1e-300)) if b.size else 0.0
