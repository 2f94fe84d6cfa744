"""
ORACLE (test infrastructure, see oracle/__init__.py) — floating-point side of the hot path, torch CPU.

Reference-shaped formulation on purpose (per-pair gathered parameters, nested forward-mode jvp, index-add
segment sums, reverse mode over everything): it is the restatement of what the reference does, not of
how the CUDA kernels do it.

Restates, from the reference's behaviour:
  * get_jmaps                                   fbpinns/trainers.py:63-107
  * FBPINN_model_inner / FBPINN_model           fbpinns/trainers.py:113-118, 126-177
  * FBPINN_forward / _get_ujs / jacfwd          fbpinns/trainers.py:197-203, 213-247
  * FBPINN_loss                                 fbpinns/trainers.py:249-267
  * FBPINN_update                               fbpinns/trainers.py:285-296  (+ optax Adam, see ref_adam)
  * norm_fn / unnorm_fn / window_fn             fbpinns/decompositions.py:183-199
  * windows.cosine                              fbpinns/windows.py:25-35
  * networks.FCN.network_fn, norm, unnorm       fbpinns/networks.py:61-68, 197-201

Parameter container here (plain python, torch leaves):
  decomp : dict as returned by oracle.ref_takes.*_init_params with leaves converted by `to_torch`
  layers : list of (W (m, out, in), b (m, out)) — the reference's "subdomain" leaves
"""

import math

import numpy as np
import torch
from torch.func import jvp


# --------------------------------------------------------------------------- jmaps

def get_jmaps(required_ujs):
    """fbpinns/trainers.py:63-107.  Returns (nodes, leaves, jac_is) with the reference's exact contents."""
    trie = {}
    for _, path in required_ujs:
        t = trie
        for ix in path:
            t = t.setdefault(ix, {})

    nodes = []

    def walk(t, parent_fn, prefix):
        for ix, child in t.items():
            path = prefix + (ix,)
            nodes.append(((parent_fn, ix), path, 0 if child else 1))
            if child:
                walk(child, len(nodes), path)       # fs index of the node just appended
    walk(trie, 0, ())
    nodes = tuple(nodes)

    leaves = tuple((i + 1, nd[1]) for i, nd in enumerate(nodes) if nd[2])
    if not leaves:
        leaves = ((0, ()),)

    jac_is = ()
    for iu, path in required_ujs:
        io = len(path)
        il = [lf[1][:io] for lf in leaves].index(tuple(path))
        jac_is += ((il, io, iu),)
    return nodes, leaves, jac_is


# --------------------------------------------------------------------------- single-pair maths, batched over pairs

def cosine_window(xmin, xmax, x):
    """fbpinns/windows.py:25-35, batched over the leading (pair) axis; returns (s, 1)."""
    mu, sd = (xmin + xmax) / 2, (xmax - xmin) / 2
    ws = ((1 + torch.cos(math.pi * (x - mu) / sd)) / 2) ** 2
    step = lambda a: (a >= 0).to(x.dtype)          # jnp.heaviside(a, 1): zero derivative
    ws = step(x - xmin) * step(xmax - x) * ws
    return torch.prod(ws, dim=1, keepdim=True)


def network_forward(network, layers_take, h, static_take=None):
    """Network.network_fn of the reference, batched over the leading (pair) axis: every leaf of `layers_take` carries the
    pair's own parameters.  network: "fcn" (fbpinns/networks.py:61-68), "adaptive_fcn" (:93-101), "siren" (:125-133),
    "adaptive_siren" (:158-166), "fourier" (:183-194, static_take = {"omega": (s, n_features, xd)})."""
    mv = lambda w, v: torch.einsum("soi,si->so", w, v)
    if network == "fourier":
        h = mv(static_take["omega"], h)
        h = torch.cat([torch.sin(h), torch.cos(h)], dim=1)
        network = "fcn"
    if network == "fcn":
        for w, b in layers_take[:-1]:
            h = torch.tanh(mv(w, h) + b)
    elif network == "adaptive_fcn":
        for w, b, a in layers_take[:-1]:
            h = a * torch.tanh((mv(w, h) + b) / a)
    elif network == "siren":
        for w, b in layers_take[:-1]:
            h = torch.sin(mv(w, h) + b)
    elif network == "adaptive_siren":
        for w, b, c, o in layers_take[:-1]:
            h = c * torch.sin(o * (mv(w, h) + b))
    else:
        raise ValueError(f"unknown network {network}")
    w, b = layers_take[-1][0], layers_take[-1][1]
    return mv(w, h) + b


def model_inner(ps_take, layers_take, x_take, network="fcn", static_take=None):
    """FBPINN_model_inner (fbpinns/trainers.py:113-118) for a batch of pairs.
    ps_take = [xmins, xmaxs, wmins, wmaxs, flags, unnorms] gathered per pair."""
    xmin, xmax = ps_take[0], ps_take[1]
    mu, sd = (xmax + xmin) / 2, (xmax - xmin) / 2
    h = (x_take - mu) / sd                                           # norm_fn, decompositions.py:183-188
    u_raw = network_forward(network, layers_take, h, static_take)   # Network.network_fn
    un = ps_take[5]
    u = u_raw * un[:, 1:2] + un[:, 0:1]                             # unnorm_fn, decompositions.py:190-194
    flag = ps_take[4]
    win = flag * cosine_window(xmin, xmax, x_take) + (1 - flag)     # window_fn, decompositions.py:196-199
    return u * win, win, u_raw


def fbpinn_model(decomp_cut, layers_cut, x_batch, takes, constraining_fn=None, all_params=None, network="fcn",
                 net_static_cut=None):
    """FBPINN_model (fbpinns/trainers.py:126-177).  decomp_cut / layers_cut are already cut to all_ims
    order (static_params = cut_all(...), trainable = concat(active, fixed))."""
    m_take, n_take, p_take, np_take, npou = takes
    m_take, n_take, p_take, np_take = [torch.as_tensor(np.asarray(t), dtype=torch.long)
                                       for t in (m_take, n_take, p_take, np_take)]
    x_take = x_batch[n_take]
    ps_take = [p[m_take] for p in decomp_cut["subdomain"]["params"]]
    layers_take = [tuple(t[m_take] for t in leaf) for leaf in layers_cut]
    static_take = None if net_static_cut is None else {k: v[m_take] for k, v in net_static_cut.items()}
    us, ws, us_raw = model_inner(ps_take, layers_take, x_take, network, static_take)

    cat = torch.cat([us, ws], dim=1)
    seg = torch.zeros((len(np_take), cat.shape[1]), dtype=cat.dtype).index_add(0, p_take, cat)
    wp = seg[:, -1:]
    u = seg[:, :-1] / wp
    u = torch.zeros((x_batch.shape[0], u.shape[1]), dtype=u.dtype).index_add(0, np_take, u)
    u = u / npou
    if constraining_fn is not None:
        u = constraining_fn(all_params, x_batch, u)
    return u, wp, us, ws, us_raw


def _jacfwd(f, v):
    """fbpinns/trainers.py:241-247."""
    def jacfun(x):
        y, j, aux = jvp(f, (x,), (v,), has_aux=True)
        return j, aux + (y,)
    return jacfun


def get_ujs(x_batch, jmaps, u_fn):
    """_get_ujs, fbpinns/trainers.py:213-239; u_fn(x_batch) -> (u (n, ud), ())."""
    nodes, leaves, jac_is = jmaps
    n, xd = x_batch.shape
    vs = torch.eye(xd, dtype=x_batch.dtype).repeat(n, 1, 1)
    fs = [u_fn]
    for (ni, ix), _, _ in nodes:
        fs.append(_jacfwd(fs[ni], vs[:, ix]))
    jacs = []
    for ie, _ in leaves:
        fin, jac = fs[ie](x_batch)
        jacs.append(jac + (fin,))
    return [jacs[il][io][:, iu:iu + 1] for il, io, iu in jac_is]


def fbpinn_forward(decomp_cut, layers_cut, x_batch, takes, jmaps, constraining_fn=None, all_params=None, network="fcn",
                   net_static_cut=None):
    """FBPINN_forward, fbpinns/trainers.py:197-203."""
    def u_fn(xb):
        return fbpinn_model(decomp_cut, layers_cut, xb, takes, constraining_fn, all_params, network, net_static_cut)[0], ()
    return get_ujs(x_batch, jmaps, u_fn)


def fbpinn_loss(active_layers, fixed_layers, decomp_cut, takess, constraints, jmapss, loss_fn,
                constraining_fn=None, make_all_params=None, network="fcn", net_static_cut=None):
    """FBPINN_loss, fbpinns/trainers.py:249-267.
    make_all_params(layers_cut) builds whatever `all_params` object loss_fn / constraining_fn expect."""
    layers_cut = [tuple(torch.cat([ta, tf], 0) for ta, tf in zip(la, lf)) for la, lf in zip(active_layers, fixed_layers)]
    all_params = make_all_params(layers_cut) if make_all_params is not None else None
    out = []
    for takes, jmaps, constraint in zip(takess, jmapss, constraints):
        x_batch = constraint[0]
        ujs = fbpinn_forward(decomp_cut, layers_cut, x_batch, takes, jmaps, constraining_fn, all_params, network,
                             net_static_cut)
        out.append(list(constraint) + ujs)
    return loss_fn(all_params, out)


# --------------------------------------------------------------------------- helpers

def to_torch(decomp, dtype):
    """numpy static decomposition dict → torch leaves of the given dtype."""
    d = dict(decomp)
    d["subdomain"] = {"params": [torch.as_tensor(np.asarray(p), dtype=dtype) for p in decomp["subdomain"]["params"]],
                      "pou": torch.as_tensor(np.asarray(decomp["subdomain"]["pou"]), dtype=dtype)}
    return d


def cut_decomp(decomp_t, ims):
    """cut_all on the static decomposition leaves (fbpinns/trainers.py:404-408)."""
    ims = torch.as_tensor(np.asarray(ims), dtype=torch.long)
    d = dict(decomp_t)
    d["subdomain"] = {"params": [p[ims] for p in decomp_t["subdomain"]["params"]],
                      "pou": decomp_t["subdomain"]["pou"][ims]}
    return d


def cut_layers(layers, ims):
    ims = torch.as_tensor(np.asarray(ims), dtype=torch.long)
    return [tuple(t[ims] for t in leaf) for leaf in layers]


def init_fcn_params(rng, m, layer_sizes, dtype=np.float32):
    """Synthetic parameters with the distribution of FCN._random_layer_params (fbpinns/networks.py:48-58):
    U(-1/sqrt(fan_in), +1/sqrt(fan_in)), drawn from a numpy Generator (jax.random bits are unpinned, so
    parameters are always explicit inputs to both sides)."""
    layers = []
    for fan_in, fan_out in zip(layer_sizes[:-1], layer_sizes[1:]):
        v = np.sqrt(1 / fan_in)
        w = rng.uniform(-v, v, size=(m, fan_out, fan_in)).astype(dtype)
        b = rng.uniform(-v, v, size=(m, fan_out)).astype(dtype)
        layers.append((w, b))
    return layers
