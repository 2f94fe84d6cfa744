"""
ORACLE (test infrastructure, see oracle/__init__.py) — Adam as used by the reference.

The reference calls `optax.adam(**c.optimiser_kwargs)` (fbpinns/trainers.py:430), `optimiser.update`
(:294) and `optax.apply_updates` (:295).  optax (pinned only as `optax>=0.1.4`, pyproject.toml:22) is a
third-party dependency that is NOT in the reference tree and not installed here, so this is a
restatement of its published algorithm (`scale_by_adam` followed by `scale(-learning_rate)`), defaults
b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0.  PARITY UNPINNED at the bit level (no reference test or golden
vector exercises Adam numerics); checked against closed-form few-step sequences in tests/test_adam.py.

State = (count:int32 scalar, mu, nu); `count` is ONE counter shared by the whole parameter tree — when the
active set changes the reference cuts/merges mu/nu per subdomain but keeps the first tree's non-dict leaf,
i.e. the global count (fbpinns/trainers.py:51-60, 531, 640).
"""

import numpy as np


def adam_init(params):
    return dict(count=np.int32(0), mu=[np.zeros_like(p) for p in params], nu=[np.zeros_like(p) for p in params])


def adam_update(grads, state, params, learning_rate=1e-3, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    """One optax.adam step on lists of numpy arrays (all of one float dtype). Returns new_params, new_state."""
    dt = params[0].dtype.type if params else np.float32
    count = np.int32(state["count"] + 1)
    mu = [dt(b1) * m + dt(1 - b1) * g for m, g in zip(state["mu"], grads)]
    nu = [dt(b2) * v + dt(1 - b2) * (g * g) for v, g in zip(state["nu"], grads)]
    c1 = dt(1) - dt(b1) ** dt(count)
    c2 = dt(1) - dt(b2) ** dt(count)
    new_params = []
    for p, m, v in zip(params, mu, nu):
        m_hat = m / c1
        v_hat = v / c2
        upd = m_hat / (np.sqrt(v_hat + dt(eps_root)) + dt(eps))
        new_params.append(p + dt(-learning_rate) * upd)
    return new_params, dict(count=count, mu=mu, nu=nu)
