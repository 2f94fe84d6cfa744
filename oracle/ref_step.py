"""
ORACLE (test infrastructure, see oracle/__init__.py) — one FBPINN_update step on CPU.

Restates FBPINN_update (fbpinns/trainers.py:285-296): value_and_grad of FBPINN_loss w.r.t. the ACTIVE
trainable parameters only (active subdomain network leaves + problem trainables), then optax Adam
(oracle.ref_adam) and apply_updates.  Fixed subdomains take part in the forward only.
"""

import numpy as np
import torch

from . import ref_model, ref_adam


def loss_and_grads(active_layers, fixed_layers, problem_trainable, decomp_cut, takess, constraints, jmapss,
                   loss_fn, constraining_fn, make_all_params, dtype=torch.float32):
    """Returns loss (python float), grads for active layers [(gW, gb), ...] and for problem trainables (dict)."""
    al = [(torch.tensor(np.asarray(w), dtype=dtype, requires_grad=True),
           torch.tensor(np.asarray(b), dtype=dtype, requires_grad=True)) for w, b in active_layers]
    fl = [(torch.as_tensor(np.asarray(w), dtype=dtype), torch.as_tensor(np.asarray(b), dtype=dtype))
          for w, b in fixed_layers]
    pt = {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=True) for k, v in problem_trainable.items()}
    cons = [[torch.as_tensor(np.asarray(c), dtype=dtype) for c in con] for con in constraints]
    loss = ref_model.fbpinn_loss(al, fl, decomp_cut, takess, cons, jmapss, loss_fn, constraining_fn,
                                 (lambda layers_cut: make_all_params(layers_cut, pt)))
    leaves = [t for wb in al for t in wb] + list(pt.values())
    grads = torch.autograd.grad(loss, leaves, allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(l) for g, l in zip(grads, leaves)]
    nl = 2 * len(al)
    g_layers = [(grads[2 * i].numpy(), grads[2 * i + 1].numpy()) for i in range(len(al))]
    g_prob = {k: grads[nl + i].numpy() for i, k in enumerate(pt)}
    return float(loss.detach()), g_layers, g_prob


def update(active_layers, fixed_layers, problem_trainable, opt_state, decomp_cut, takess, constraints, jmapss,
           loss_fn, constraining_fn, make_all_params, dtype=torch.float32, **adam_kwargs):
    """One FBPINN_update.  Parameters are lists of numpy arrays; opt_state from ref_adam.adam_init over the
    flat leaf list [W0, b0, W1, b1, ..., *problem_trainables]."""
    npdt = np.float32 if dtype == torch.float32 else np.float64
    loss, g_layers, g_prob = loss_and_grads(active_layers, fixed_layers, problem_trainable, decomp_cut, takess,
                                            constraints, jmapss, loss_fn, constraining_fn, make_all_params, dtype)
    flat_p = [np.asarray(t, dtype=npdt) for wb in active_layers for t in wb] + \
             [np.asarray(v, dtype=npdt) for v in problem_trainable.values()]
    flat_g = [np.asarray(t, dtype=npdt) for wb in g_layers for t in wb] + \
             [np.asarray(g_prob[k], dtype=npdt) for k in problem_trainable]
    new_p, opt_state = ref_adam.adam_update(flat_g, opt_state, flat_p, **adam_kwargs)
    nl = len(active_layers)
    new_layers = [(new_p[2 * i], new_p[2 * i + 1]) for i in range(nl)]
    new_prob = {k: new_p[2 * nl + i] for i, k in enumerate(problem_trainable)}
    return loss, new_layers, new_prob, opt_state
