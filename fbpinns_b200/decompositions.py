"""
Domain decompositions (API of fbpinns/decompositions.py:31-227, 338-375).

`init_params` reproduces the reference's float64 box arithmetic (then float32 casts) bit for bit — the grid is a
tensor product, so the per-axis 1-D edge/overlap arrays are computed first and broadcast with the same "ij"
meshgrid flattening (subdomain index = C-order index of the grid, fbpinns/decompositions.py:163).
`inside_points` / `inside_models` run on the device through the C ABI (fbp_inside_count / fbp_takes_*).
`norm_fn`, `unnorm_fn`, `window_fn` are host mirrors of what the kernels compute per pair.
"""
import numpy as np
import torch

from . import windows, networks


class Decomposition:
    """Base decomposition class (fbpinns/decompositions.py:31-94)."""

    @staticmethod
    def init_params(*args):
        raise NotImplementedError

    @staticmethod
    def norm_fn(params, x):
        raise NotImplementedError

    @staticmethod
    def unnorm_fn(params, u):
        raise NotImplementedError

    @staticmethod
    def window_fn(params, x):
        raise NotImplementedError

    @staticmethod
    def inside_points(all_params, x_batch):
        raise NotImplementedError

    @staticmethod
    def inside_models(all_params, x_batch, ims):
        raise NotImplementedError


def _axis_arrays(x, w):
    "1-D edges and overlap widths along one axis (float64)"
    x, w = np.asarray(x, dtype=np.float64), np.asarray(w, dtype=np.float64)
    lo, hi = x - w / 2, x + w / 2
    wlo, whi = 0.5 * (hi - lo), 0.5 * (hi - lo)
    if len(x) > 1:
        ov = hi[:-1] - lo[1:]
        whi[:-1] = ov
        wlo[1:] = ov
        wlo[0] = whi[0]
        whi[-1] = wlo[-1]
    return lo, hi, wlo, whi


class RectangularDecompositionND(Decomposition):
    """ND hyperrectangular domain; rectangular subdomains on a regular grid of centres/widths."""

    @staticmethod
    def init_params(subdomain_xs, subdomain_ws, unnorm):
        nm = tuple(len(x) for x in subdomain_xs)
        xd = len(subdomain_xs)
        ps = RectangularDecompositionND._get_level_params(0, xd, subdomain_xs, subdomain_ws, unnorm)
        return RectangularDecompositionND._static(ps, int(np.prod(nm)), xd), {}

    @staticmethod
    def _static(ps, m, xd):
        xmins0, xmaxs0 = ps[0] + ps[2] / 2, ps[1] - ps[3] / 2        # float64 centre lines of the overlaps
        params = [torch.tensor(np.asarray(p, dtype=np.float32)) for p in ps]
        return {"m": m, "xd": xd, "subdomain": {"params": params[:-1], "pou": params[-1]},
                "xmins0": xmins0, "xmaxs0": xmaxs0}

    @staticmethod
    def _get_level_params(il, xd, subdomain_xs, subdomain_ws, unnorm):
        if [len(x) for x in subdomain_xs] != [len(w) for w in subdomain_ws]:
            raise ValueError("shape of subdomain_ws not same as subdomain_xs")
        per_axis = [_axis_arrays(x, w) for x, w in zip(subdomain_xs, subdomain_ws)]
        cols = []
        for which in range(4):
            grids = np.meshgrid(*[pa[which] for pa in per_axis], indexing="ij")
            cols.append(np.stack([g.reshape(-1) for g in grids], axis=1))       # (m, xd)
        xmins, xmaxs, wmins, wmaxs = cols
        if (wmins <= 0).any() or (wmaxs <= 0).any():
            raise ValueError("some subdomains are not overlapping!")
        m = xmins.shape[0]
        flags = np.zeros((m, 1)) if m == 1 else np.ones((m, 1))
        unnorms = np.concatenate([unnorm[0] * np.ones((m, 1)), unnorm[1] * np.ones((m, 1))], axis=1)
        pous = il * np.ones((m, 1))
        return [xmins, xmaxs, wmins, wmaxs, flags, unnorms, pous]

    # ---- host mirrors of the per-pair functions (single point, single subdomain's params) -------------------
    @staticmethod
    def norm_fn(params, x):
        p = params["static"]["decomposition"]["subdomain"]["params"]
        xmin, xmax = p[:2]
        return networks.norm((xmax + xmin) / 2, (xmax - xmin) / 2, x)

    @staticmethod
    def unnorm_fn(params, u):
        p = params["static"]["decomposition"]["subdomain"]["params"]
        mu, sd = p[5]
        return networks.unnorm(mu, sd, u)

    @staticmethod
    def window_fn(params, x):
        p = params["static"]["decomposition"]["subdomain"]["params"]
        return p[4] * windows.cosine(p[0], p[1], x) + (1 - p[4])

    # ---- index construction on the device ---------------------------------------------------------------------
    @staticmethod
    def _device(all_params, device=None):
        from .engine import DeviceDecomposition
        d = all_params["static"]["decomposition"]
        dev = torch.device(device or "cuda:0")
        cache = d.setdefault("_device_cache", {})
        if str(dev) not in cache:
            cache[str(dev)] = DeviceDecomposition(d["subdomain"]["params"], d["subdomain"]["pou"], dev)
        return cache[str(dev)]

    @staticmethod
    def inside_points(all_params, x_batch):
        """Returns n_take, m_take (pairs sorted by point then model, global model indices) and inside_ims
        (fbpinns/decompositions.py:201-208), as int32 CUDA tensors."""
        from .engine import DeviceTakes
        dd = RectangularDecompositionND._device(all_params, x_batch.device if x_batch.is_cuda else None)
        x = x_batch.to(dd.device, torch.float32).contiguous()
        m = dd.m
        ident = np.arange(m, dtype=np.int32)
        t = DeviceTakes(dd, x, ident, ident, m, tile_points=128, target_items=1)
        counts = torch.diff(t.sub_off)
        return t.n_take, t.m_take, torch.nonzero(counts > 0).reshape(-1).to(torch.int32)

    @staticmethod
    def inside_models(all_params, x_batch, ims):
        """Returns inside_ips (points inside >= 1 model of ims) and d (fbpinns/decompositions.py:210-215)."""
        from .engine import nonzero_i32
        dd = RectangularDecompositionND._device(all_params, x_batch.device if x_batch.is_cuda else None)
        x = x_batch.to(dd.device, torch.float32).contiguous()
        ims_d = torch.as_tensor(np.asarray(ims.cpu() if torch.is_tensor(ims) else ims, dtype=np.int32),
                                dtype=torch.int32, device=dd.device)
        pt, mc = dd.inside_count(x, models=ims_d)
        d = float(mc.double().mean().item() ** (1 / x.shape[1])) if ims_d.numel() else float("nan")
        return nonzero_i32(pt), d


class MultilevelRectangularDecompositionND(RectangularDecompositionND):
    """Several rectangular decompositions at different scales, one partition of unity per level
    (fbpinns/decompositions.py:338-375)."""

    @staticmethod
    def init_params(subdomain_xss, subdomain_wss, unnorm):
        nms = [tuple(len(x) for x in sx) for sx in subdomain_xss]
        if False in [len(nm) == len(nms[0]) for nm in nms]:
            raise ValueError("subdomain_xss are not all the same dimensionality")
        xd = len(subdomain_xss[0])
        cols = [[] for _ in range(7)]
        for il, (sx, sw) in enumerate(zip(subdomain_xss, subdomain_wss)):
            for i, p in enumerate(RectangularDecompositionND._get_level_params(il, xd, sx, sw, unnorm)):
                cols[i].append(p)
        ps = [np.concatenate(c) for c in cols]
        return RectangularDecompositionND._static(ps, int(sum(np.prod(nm) for nm in nms)), xd), {}
