"""
Checkpoints in the reference's on-disk format (fbpinns/trainers_base.py:64-69, fbpinns/trainers.py:721,
fbpinns/analysis.py:18-50): `model_{i:08d}.jax` is a pickle of

    (i, all_params, all_opt_states, active, u_test_losses)

with every array a numpy array, `all_opt_states` = the state of `optax.adam`:
`(ScaleByAdamState(count, mu, nu), EmptyState())`, mu / nu mirroring `all_params["trainable"]`.

optax is not installed here, but pickle stores namedtuples by reference to their defining module, so writing (and
reading) such a file needs *a* class at `optax._src.transform.ScaleByAdamState` / `optax._src.base.EmptyState`.  The
context manager below registers field-compatible stand-ins under those names when the real package is absent and
removes them again; a file written that way unpickles into the real optax classes wherever optax exists, and a file
written by the reference unpickles here into the stand-ins.  (Adam arithmetic itself stays "parity unpinned".)
"""
import contextlib
import pickle
import sys
import types
from collections import namedtuple

import numpy as np


def _to_numpy(tree):
    "torch tensors / numpy arrays -> numpy, containers preserved (tuples stay tuples), private '_...' keys dropped"
    try:
        import torch
    except ImportError:          # pragma: no cover
        torch = None
    if torch is not None and torch.is_tensor(tree):
        return tree.detach().cpu().numpy()
    if isinstance(tree, dict):
        return {k: _to_numpy(v) for k, v in tree.items() if not (isinstance(k, str) and k.startswith("_"))}
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
        return type(tree)(*[_to_numpy(v) for v in tree])
    if isinstance(tree, (list, tuple)):
        return type(tree)(_to_numpy(v) for v in tree)
    return tree


@contextlib.contextmanager
def optax_classes():
    """Yields (ScaleByAdamState, EmptyState): the real optax classes if importable, else stand-ins registered under the
    same module paths for the duration of the block."""
    try:
        from optax._src.transform import ScaleByAdamState      # noqa: F401
        from optax._src.base import EmptyState                  # noqa: F401
        yield ScaleByAdamState, EmptyState
        return
    except ImportError:
        pass
    added = []
    for name in ("optax", "optax._src", "optax._src.transform", "optax._src.base"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
            added.append(name)
    tr, base = sys.modules["optax._src.transform"], sys.modules["optax._src.base"]
    sbas = namedtuple("ScaleByAdamState", ["count", "mu", "nu"])
    sbas.__module__, sbas.__qualname__ = "optax._src.transform", "ScaleByAdamState"
    empty = namedtuple("EmptyState", [])
    empty.__module__, empty.__qualname__ = "optax._src.base", "EmptyState"
    tr.ScaleByAdamState, base.EmptyState = sbas, empty
    try:
        yield sbas, empty
    finally:
        for name in added:
            sys.modules.pop(name, None)


def save_model(path, i, all_params, mu, nu, count, active, u_test_losses):
    """Write a reference-format checkpoint.  mu / nu: trees with the structure of all_params["trainable"]; count: int."""
    with optax_classes() as (ScaleByAdamState, EmptyState):
        state = (ScaleByAdamState(count=np.asarray(count, dtype=np.int32), mu=_to_numpy(mu), nu=_to_numpy(nu)), EmptyState())
        model = (int(i), _to_numpy(all_params), state, np.asarray(active), np.asarray(u_test_losses))
        with open(path, "wb") as f:
            pickle.dump(model, f)


def load_model(path):
    """Read a checkpoint written by the reference or by save_model -> (i, all_params, (count, mu, nu), active, losses)."""
    with optax_classes():
        with open(path, "rb") as f:
            model = pickle.load(f)
    i, all_params, opt_states = model[0], model[1], model[2]
    active, losses = (model[3], model[4]) if len(model) == 5 else (None, model[3])      # PINN checkpoints carry no active set
    adam = opt_states[0]
    return int(i), all_params, (int(np.asarray(adam.count)), adam.mu, adam.nu), active, losses
