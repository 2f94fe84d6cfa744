"""
Problem domains (API of fbpinns/domains.py:18-142).  Sampling runs once at initialisation and is an INPUT of the
hot path, so it stays host code; results are float32 torch tensors.
"""
import numpy as np
import scipy.stats
import torch

from .util import jax_prng


class Domain:
    """Base domain class (fbpinns/domains.py:18-55)."""

    @staticmethod
    def init_params(*args):
        raise NotImplementedError

    @staticmethod
    def sample_interior(all_params, key, sampler, batch_shape):
        raise NotImplementedError

    @staticmethod
    def sample_boundaries(all_params, key, sampler, batch_shapes):
        raise NotImplementedError

    @staticmethod
    def norm_fn(all_params, x):
        raise NotImplementedError


def _as_np(a):
    return a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)


class RectangularDomainND(Domain):

    @staticmethod
    def init_params(xmin, xmax):
        xmin, xmax = np.asarray(xmin), np.asarray(xmax)
        assert xmin.shape == xmax.shape
        assert xmin.ndim == 1
        return {"xd": len(xmin), "xmin": torch.tensor(xmin, dtype=torch.float32),
                "xmax": torch.tensor(xmax, dtype=torch.float32)}, {}

    @staticmethod
    def sample_interior(all_params, key, sampler, batch_shape):
        d = all_params["static"]["domain"]
        return RectangularDomainND._rectangle_samplerND(key, sampler, d["xmin"], d["xmax"], batch_shape)

    @staticmethod
    def sample_boundaries(all_params, key, sampler, batch_shapes):
        d = all_params["static"]["domain"]
        xmin, xmax, xd = _as_np(d["xmin"]), _as_np(d["xmax"]), d["xd"]
        assert len(batch_shapes) == 2 * xd
        # one independent key per face.  jax-style keys follow the reference's chain `key, subkey = split(key)` per face
        # (fbpinns/domains.py:95); an int seed is expanded into per-face generators; a numpy Generator advances by itself
        keys = []
        if jax_prng.is_key(key):
            for _ in range(2 * xd):
                key, subkey = jax_prng.split(key)
                keys.append(subkey)
        elif isinstance(key, np.random.Generator) or key is None:
            keys = [key] * (2 * xd)                 # None only occurs with the grid / quasi-random samplers
        else:
            keys = list(np.random.default_rng(key).spawn(2 * xd))
        out = []
        for i in range(xd):
            ic = [j for j in range(xd) if j != i]
            for j, v in enumerate([xmin[i], xmax[i]]):
                bs = batch_shapes[2 * i + j]
                if ic:
                    xb_ = RectangularDomainND._rectangle_samplerND(keys[2 * i + j], sampler, xmin[ic], xmax[ic], bs)
                    xb = torch.full((int(np.prod(bs)), xd), float(v), dtype=torch.float32)
                    xb[:, ic] = xb_
                else:
                    assert len(bs) == 1
                    xb = torch.full(tuple(bs) + (1,), float(v), dtype=torch.float32)
                out.append(xb)
        return out

    @staticmethod
    def norm_fn(all_params, x):
        d = all_params["static"]["domain"]
        mu, sd = (d["xmax"] + d["xmin"]) / 2, (d["xmax"] - d["xmin"]) / 2
        return (x - mu.to(x.device)) / sd.to(x.device)

    @staticmethod
    def _rectangle_samplerND(key, sampler, xmin, xmax, batch_shape):
        "Flattened samples of x in a rectangle, on a mesh or (quasi-)random (fbpinns/domains.py:111-142)"
        xmin, xmax = _as_np(xmin).astype(np.float64), _as_np(xmax).astype(np.float64)
        assert xmin.shape == xmax.shape
        assert xmin.ndim == 1
        xd = len(xmin)
        assert len(batch_shape) == xd
        if sampler not in ["grid", "uniform", "sobol", "halton"]:
            raise ValueError("ERROR: unexpected sampler")
        if sampler == "grid":
            xs = [np.linspace(a, b, n) for a, b, n in zip(xmin, xmax, batch_shape)]
            x_batch = np.stack(np.meshgrid(*xs, indexing="ij"), -1).reshape(-1, xd)
        else:
            n = int(np.prod(batch_shape))
            if sampler == "halton":
                s = scipy.stats.qmc.Halton(xd).random(n)
            elif sampler == "sobol":
                s = scipy.stats.qmc.Sobol(xd).random(n)
            elif jax_prng.is_key(key):
                s = jax_prng.uniform(key, (n, xd)).astype(np.float64)      # jax.random.uniform(key, (n, xd)), restated
            else:
                rng = key if isinstance(key, np.random.Generator) else np.random.default_rng(key)
                s = rng.random((n, xd))
            x_batch = xmin.reshape(1, -1) + (xmax - xmin).reshape(1, -1) * s
        return torch.tensor(x_batch, dtype=torch.float32)
