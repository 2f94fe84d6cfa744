"""Accuracy probe of the reverse kernels at the FULL cfg 5 size against a float64 reference (the oracle's per-pair model
and nested-jvp jets, evaluated on the CPU for a sample of subdomains with all their pairs):
    L_sub = sum_{pairs of the subdomain} grow[row(pair)] . jets(u w)(pair),   gradient w.r.t. the subdomain's parameters.
Reports, per parameter group and kernel (tiled FFMA2 / tensor reverse 1 / tensor reverse 2), the error relative to the
group's largest gradient entry over the sampled subdomains and relative to each subdomain's own largest entry.
Writes gpurun_out/bwd_precision.json.      python tests/tools/bwd_precision.py [n_subdomains]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fbpinns_b200 import configs                                        # noqa: E402
from fbpinns_b200.engine import Plan, ConstraintEvaluator, unpack_params  # noqa: E402
from fbpinns_b200.trainers import FBPINNTrainer                         # noqa: E402
from oracle import ref_model                                            # noqa: E402


def main():
    nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    kw = {} if "--small" not in sys.argv else dict(n_sub=(8, 8), n_pts=(128, 128))
    c = configs.cfg5_poisson(device="cuda:0", kernel="tiled", use_cuda_graph=False, **kw)
    tr = FBPINNTrainer(c)
    tr.setup()
    m = tr.all_params["static"]["decomposition"]["m"]
    tr.set_active(np.ones(m, dtype=int))
    ev = tr.inputs.evaluators[0]
    ev.set_affine(None)
    jet, takes = ev.plan.jet, ev.takes
    torch.manual_seed(0)
    ubar = torch.randn(takes.n, ev.V, device="cuda")
    grads = {}
    variants = [("tiled", "tiled", "1"), ("tensor1", "tensor-full", "1"), ("tensor2", "tensor-full", "2")]
    for tag, kernel, bwd in variants:
        os.environ["FBP_TC_BWD"] = bwd
        plan = Plan(ev.plan.layer_sizes, jet, kernel=kernel)
        e = ConstraintEvaluator(plan, takes, ev.x, tr.dd) if kernel != "tiled" else ev
        g = torch.full((takes.m_active, tr.params.shape[1]), float("nan"), device="cuda")
        e.forward(tr.params)
        e.backward(ubar, tr.params, g, accumulate=False)
        torch.cuda.synchronize()
        grads[tag] = [(w.cpu().double().numpy(), b.cpu().double().numpy()) for w, b in unpack_params(ev.plan, g)]
    grow = ev.grow.cpu().double()
    sub_off = takes.sub_off.cpu().numpy()
    sp_point, sp_row = takes.spair_point.cpu().numpy(), takes.spair_row.cpu().numpy()
    x = ev.x.cpu().double()
    sub_ids = takes.sub_ids.cpu().numpy()
    layers = [(w.cpu().double(), b.cpu().double()) for w, b in unpack_params(ev.plan, tr.params)]
    dparams = [torch.as_tensor(np.asarray(p.cpu() if torch.is_tensor(p) else p), dtype=torch.float64)
               for p in tr.all_params["static"]["decomposition"]["subdomain"]["params"]]
    jmaps = ref_model.get_jmaps(tuple((0, p) for p in jet.comps))
    # sample: the subdomains with the largest tiled gradient + corners/edges + random interior ones
    gmax = np.max(np.abs(grads["tiled"][1][0]).reshape(takes.m_active, -1), axis=1)
    rng = np.random.default_rng(0)
    pick = list(np.argsort(-gmax)[:max(2, nsub // 3)]) + [0, takes.m_active - 1] + list(rng.integers(0, takes.m_active, nsub))
    pick = list(dict.fromkeys(int(p) for p in pick))[:nsub]
    ref = {}
    for sp in pick:
        im = int(sub_ids[sp])
        a, b = int(sub_off[sp]), int(sub_off[sp + 1])
        xs = x[sp_point[a:b]]
        rows = torch.as_tensor(sp_row[a:b], dtype=torch.long)
        leaves = [(w[im].clone().requires_grad_(True), bb[im].clone().requires_grad_(True)) for w, bb in layers]
        s = b - a
        ps_take = [p[im].expand(s, *p.shape[1:]) for p in dparams]
        lay_take = [(w.expand(s, *w.shape), bb.expand(s, *bb.shape)) for w, bb in leaves]

        def u_fn(xb):
            return ref_model.model_inner(ps_take, lay_take, xb)[0], ()
        jets = torch.cat(ref_model.get_ujs(xs, jmaps, u_fn), dim=1)          # (s, C) in jet.comps order
        L = (grow[rows] * jets).sum()
        gr = torch.autograd.grad(L, [t for wb in leaves for t in wb])
        ref[sp] = [(gr[2 * i].numpy(), gr[2 * i + 1].numpy()) for i in range(len(leaves))]
    names = ["W0", "b0", "W1", "b1", "W2", "b2"]
    res = {"pairs": int(takes.s), "subdomains": [int(p) for p in pick]}
    for tag, _, _ in variants:
        for l in range(len(layers)):
            for which in (0, 1):
                nm = names[2 * l + which]
                gscale = max(np.max(np.abs(ref[sp][l][which])) for sp in pick)
                e_glob = max(np.max(np.abs(grads[tag][l][which][sp] - ref[sp][l][which])) for sp in pick) / gscale
                e_loc = max(np.max(np.abs(grads[tag][l][which][sp] - ref[sp][l][which])) / np.max(np.abs(ref[sp][l][which])) for sp in pick)
                res[f"{tag}_{nm}_vs_group_max"] = float(e_glob)
                res[f"{tag}_{nm}_vs_own_max"] = float(e_loc)
    print(json.dumps(res, indent=1))
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    json.dump(res, open(os.path.join(d, "bwd_precision.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
