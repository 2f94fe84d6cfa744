"""
ORACLE (test infrastructure, see oracle/__init__.py) — integer / index side of the hot path, numpy.

Restates, from the reference's behaviour:
  * RectangularDecompositionND.init_params / _get_level_params     fbpinns/decompositions.py:105-181
  * MultilevelRectangularDecompositionND.init_params               fbpinns/decompositions.py:342-375
  * _inside_rectangleND / inside_points / inside_models            fbpinns/decompositions.py:201-227
    (the dense `nonzero` of the (n, m) inside mask is the specification:
     fbpinns/decompositions_base.py:100-106; :23-82 is a memory-bounded evaluation of it)
  * get_inputs                                                     fbpinns/trainers.py:332-419
  * FBPINNTrainer._get_x_batch / _get_update_inputs                fbpinns/trainers.py:482-575
  * _common_train_initialisation (constraint bookkeeping part)     fbpinns/trainers.py:443-450

dtypes follow JAX defaults with x64 disabled: float32 values, int32 indices.
"""

import numpy as np

I32 = np.int32
F32 = np.float32


# --------------------------------------------------------------------------- decomposition (A1)

def level_params(il, xd, subdomain_xs, subdomain_ws, unnorm):
    """fbpinns/decompositions.py:135-181.  float64 numpy in, list of 7 float64 (m, .) arrays out:
    [xmins, xmaxs, wmins, wmaxs, flags, unnorms, pous]; subdomain index = C-order index of the grid."""
    xs = np.stack(np.meshgrid(*subdomain_xs, indexing="ij"), 0)
    ws = np.stack(np.meshgrid(*subdomain_ws, indexing="ij"), 0)
    if xs.shape != ws.shape:
        raise ValueError("shape of subdomain_ws not same as subdomain_xs")
    lo, hi = xs - ws / 2, xs + ws / 2

    # overlap widths: default (single subdomain along an axis) is half the box width
    wlo, whi = 0.5 * (hi - lo), 0.5 * (hi - lo)
    for ax in range(xd):
        inner_l = [slice(None)] * xd
        inner_r = [slice(None)] * xd
        inner_l[ax] = slice(None, -1)
        inner_r[ax] = slice(1, None)
        a, b = (ax,) + tuple(inner_l), (ax,) + tuple(inner_r)
        ov = hi[a] - lo[b]          # overlap between neighbour i (right edge) and i+1 (left edge)
        whi[a] = ov
        wlo[b] = ov
        first = [slice(None)] * xd
        last = [slice(None)] * xd
        first[ax] = 0
        last[ax] = -1
        f, l = (ax,) + tuple(first), (ax,) + tuple(last)
        wlo[f] = whi[f]
        whi[l] = wlo[l]
    if (wlo <= 0).any() or (whi <= 0).any():
        raise ValueError("some subdomains are not overlapping!")

    flat = lambda a: a.reshape(xd, -1).T
    xmins, xmaxs, wmins, wmaxs = flat(lo), flat(hi), flat(wlo), flat(whi)
    m = xmins.shape[0]
    flags = np.zeros((m, 1)) if m == 1 else np.ones((m, 1))
    unnorms = np.concatenate([unnorm[0] * np.ones((m, 1)), unnorm[1] * np.ones((m, 1))], axis=1)
    pous = il * np.ones((m, 1))
    return [xmins, xmaxs, wmins, wmaxs, flags, unnorms, pous]


def _pack_static(ps, m, xd):
    xmins0, xmaxs0 = ps[0] + ps[2] / 2, ps[1] - ps[3] / 2          # float64, for the schedulers
    params32 = [np.asarray(p, dtype=F32) for p in ps]             # jnp.array(x): f64 -> f32
    return {"m": int(m), "xd": int(xd),
            "subdomain": {"params": params32[:-1], "pou": params32[-1]},
            "xmins0": xmins0, "xmaxs0": xmaxs0}


def rectangular_init_params(subdomain_xs, subdomain_ws, unnorm):
    """fbpinns/decompositions.py:105-133 → static_params dict (numpy leaves)."""
    nm = tuple(len(x) for x in subdomain_xs)
    xd = len(subdomain_xs)
    ps = level_params(0, xd, subdomain_xs, subdomain_ws, unnorm)
    return _pack_static(ps, int(np.prod(nm)), xd)


def multilevel_init_params(subdomain_xss, subdomain_wss, unnorm):
    """fbpinns/decompositions.py:342-375."""
    nms = [tuple(len(x) for x in sx) for sx in subdomain_xss]
    if False in [len(nm) == len(nms[0]) for nm in nms]:
        raise ValueError("subdomain_xss are not all the same dimensionality")
    xd = len(subdomain_xss[0])
    cols = [[] for _ in range(7)]
    for il, (sx, sw) in enumerate(zip(subdomain_xss, subdomain_wss)):
        for i, p in enumerate(level_params(il, xd, sx, sw, unnorm)):
            cols[i].append(p)
    ps = [np.concatenate(c) for c in cols]
    return _pack_static(ps, int(sum(np.prod(nm) for nm in nms)), xd)


# --------------------------------------------------------------------------- inside tests (A2)

def inside_mask(decomp, x_batch, ims):
    """fbpinns/decompositions.py:217-227: all_d( x >= xmin & x <= xmax ) in float32 → (n, mc) bool."""
    ps = decomp["subdomain"]["params"]
    x = np.asarray(x_batch, dtype=F32)[:, None, :]
    lo = ps[0][ims][None]
    hi = ps[1][ims][None]
    return np.all((x >= lo) & (x <= hi), axis=-1)


def _batched(decomp, x_batch, ims, batch=65536):
    n = x_batch.shape[0]
    for i0 in range(0, n, batch):
        yield i0, inside_mask(decomp, x_batch[i0:i0 + batch], ims)


def inside_points(decomp, x_batch):
    """fbpinns/decompositions.py:201-208 + decompositions_base.py:70-76.
    Returns n_take, m_take (row-major nonzeros of the (n, m) mask) and inside_ims."""
    m = decomp["m"]
    ims = np.arange(m)
    n_parts, m_parts = [], []
    any_m = np.zeros(m, dtype=bool)
    for i0, msk in _batched(decomp, x_batch, ims):
        nt, mt = np.nonzero(msk)
        n_parts.append(nt + i0)
        m_parts.append(mt)
        any_m |= msk.any(0)
    n_take = np.concatenate(n_parts).astype(I32) if n_parts else np.zeros(0, I32)
    m_take = np.concatenate(m_parts).astype(I32) if m_parts else np.zeros(0, I32)
    return n_take, m_take, np.nonzero(any_m)[0].astype(I32)


def inside_models(decomp, x_batch, ims):
    """fbpinns/decompositions.py:210-215 + decompositions_base.py:78-82.
    Returns inside_ips (points inside >=1 model of ims) and d (log-only statistic)."""
    ims = np.asarray(ims)
    any_p = np.zeros(x_batch.shape[0], dtype=bool)
    per_m = np.zeros(len(ims), dtype=np.int64)
    for i0, msk in _batched(decomp, x_batch, ims):
        any_p[i0:i0 + msk.shape[0]] = msk.any(1)
        per_m += msk.sum(0)
    d = float(per_m.mean() ** (1 / x_batch.shape[1])) if len(ims) else float("nan")
    return np.nonzero(any_p)[0].astype(I32), d


# --------------------------------------------------------------------------- get_inputs (A3)

def get_inputs(x_batch, active, decomp):
    """fbpinns/trainers.py:332-419 (without the param-tree closures, which are pure indexing by
    active_ims / fixed_ims / all_ims).  Returns takes, all_ims, active_ims, fixed_ims, active."""
    m = decomp["m"]
    n_take, m_take, training_ims = inside_points(decomp, x_batch)

    active = np.array(active).copy()
    assert np.isin(active, [0, 1, 2]).all()
    assert active.shape == (m,)
    active[active == 0] = 1                     # inactive models still train if they hold points
    mask = np.zeros_like(active)
    mask[training_ims] = 1
    active = active * mask                      # models without any training point are discarded
    ims_ = np.arange(m)
    active_ims = ims_[active == 1]
    fixed_ims = ims_[active == 2]
    all_ims = np.concatenate([active_ims, fixed_ims])

    inv = np.zeros(m, dtype=np.int64)
    inv[all_ims] = np.arange(len(all_ims))
    m_take = inv[m_take]

    pous = decomp["subdomain"]["pou"][all_ims].astype(np.int64)
    npairs = np.stack([n_take, pous[m_take, 0]], axis=-1).astype(np.int64)
    if len(npairs):
        npu, p_take = np.unique(npairs, axis=0, return_inverse=True)
    else:
        npu, p_take = np.zeros((0, 2), np.int64), np.zeros(0, np.int64)
    p_take = p_take.reshape(-1)
    np_take = npu[:, 0]
    npou = len(np.unique(decomp["subdomain"]["pou"].astype(np.int64)))

    takes = (m_take.astype(I32), n_take.astype(I32), p_take.astype(I32), np_take.astype(I32), int(npou))
    return takes, all_ims.astype(I32), active_ims.astype(I32), fixed_ims.astype(I32), active


# --------------------------------------------------------------------------- constraints bookkeeping

def constraint_tables(constraint_sizes):
    """fbpinns/trainers.py:443-448: offsets (c,) and boolean membership (n, c)."""
    sizes = list(constraint_sizes)
    offsets = np.cumsum([0] + sizes[:-1]).astype(np.int64)
    fs = np.zeros((int(sum(sizes)), len(sizes)), dtype=bool)
    for ic, (o, s) in enumerate(zip(offsets, sizes)):
        fs[o:o + s, ic] = True
    return offsets, fs


def get_x_batch(active, decomp, x_batch_global, constraint_arrays_global, constraint_fs_global,
                constraint_offsets_global):
    """FBPINNTrainer._get_x_batch, fbpinns/trainers.py:482-507.
    constraint_arrays_global[ic] = list of per-point arrays of constraint ic (first is its x_batch)."""
    ims = np.arange(decomp["m"])[np.asarray(active) == 1]
    training_ips, d = inside_models(decomp, x_batch_global, ims)
    x_batch = x_batch_global[training_ips]
    constraint_fs = constraint_fs_global[training_ips]
    ix_ = np.arange(x_batch.shape[0])
    constraint_ips = [ix_[f] for f in constraint_fs.T]
    constraints = [[np.asarray(c_)[training_ips[constraint_ips[ic]] - constraint_offsets_global[ic]]
                    for c_ in constraint_arrays_global[ic]]
                   for ic in range(len(constraint_arrays_global))]
    return x_batch, constraints, constraint_fs, constraint_ips, training_ips, d


def split_takes(takes, n_x_batch, constraint_fs, constraint_ips):
    """Per-constraint split of the takes, fbpinns/trainers.py:544-571."""
    m_take, n_take, p_take, np_take, npou = takes
    takess = []
    iu_ = np.arange(np_take.shape[0])
    for f, ips in zip(constraint_fs.T, constraint_ips):
        f1 = f
        f2 = f[np_take]
        ius = iu_[f2]
        f3 = f1[n_take]
        f4 = f2[p_take]
        inv = np.zeros(n_x_batch, dtype=np.int64)
        inv[ips] = np.arange(len(ips))
        inv2 = np.zeros(np_take.shape[0], dtype=np.int64)
        inv2[ius] = np.arange(len(ius))
        takess.append((m_take[f3].astype(I32), inv[n_take[f3]].astype(I32),
                       inv2[p_take[f4]].astype(I32), inv[np_take[f2]].astype(I32), npou))
    return takess


def get_update_inputs(active, decomp, x_batch_global, constraint_arrays_global, constraint_fs_global,
                      constraint_offsets_global):
    """FBPINNTrainer._get_update_inputs, fbpinns/trainers.py:509-575 (index outputs only).
    Returns dict(active, active_ims, fixed_ims, all_ims, takess, constraints, x_batch, training_ips)."""
    active = np.array(active).copy()
    assert np.isin(active, [0, 1, 2]).all()
    assert active.shape == (decomp["m"],)
    x_batch, constraints, constraint_fs, constraint_ips, training_ips, d = get_x_batch(
        active, decomp, x_batch_global, constraint_arrays_global, constraint_fs_global,
        constraint_offsets_global)
    takes, all_ims, active_ims, fixed_ims, active2 = get_inputs(x_batch, active, decomp)
    takess = split_takes(takes, x_batch.shape[0], constraint_fs, constraint_ips)
    return dict(active=active2, active_ims=active_ims, fixed_ims=fixed_ims, all_ims=all_ims,
                takes=takes, takess=takess, constraints=constraints, x_batch=x_batch,
                training_ips=training_ips, constraint_ips=constraint_ips, d=d)
