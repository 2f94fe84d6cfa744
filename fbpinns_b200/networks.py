"""
Network plug-ins (API of fbpinns/networks.py:14-68, 197-201).  The kernels implement FCN with tanh; the other
reference networks (AdaptiveFCN, SIREN, AdaptiveSIREN, FourierFCN) are not on the hot path of any BASELINE
config and raise NotImplementedError in the trainer.
"""
import numpy as np
import torch


class Network:
    """Base neural network class (fbpinns/networks.py:14-34)."""

    @staticmethod
    def init_params(*args):
        raise NotImplementedError

    @staticmethod
    def network_fn(params, x):
        raise NotImplementedError


def _rng(key):
    return key if isinstance(key, np.random.Generator) else np.random.default_rng(key)


from .util import jax_prng      # noqa: E402  numpy restatement of jax.random's threefry keys


class FCN(Network):
    "Fully connected network"

    @staticmethod
    def init_params(key, layer_sizes):
        """`key` is a numpy Generator, an integer seed, or a jax-style threefry key (uint32[2], util.jax_prng.PRNGKey):
        with the latter the draws follow the reference's `random.split` / `random.uniform` sequence
        (fbpinns/networks.py:41-58) as restated in util/jax_prng.py (pinned by known-answer vectors only)."""
        if jax_prng.is_key(key):
            keys = jax_prng.split(key, len(layer_sizes) - 1)
            params = [FCN._random_layer_params(k, m, n) for k, m, n in zip(keys, layer_sizes[:-1], layer_sizes[1:])]
            return {}, {"layers": params}
        rng = _rng(key)
        params = [FCN._random_layer_params(rng, m, n) for m, n in zip(layer_sizes[:-1], layer_sizes[1:])]
        return {}, {"layers": params}

    @staticmethod
    def _random_layer_params(key, m, n):
        "U(-1/sqrt(fan_in), 1/sqrt(fan_in)) weights (n, m) and biases (n,) — fbpinns/networks.py:48-58"
        if jax_prng.is_key(key):
            w_key, b_key = jax_prng.split(key)
            v = np.float32(np.sqrt(np.float32(1 / m)))
            return (torch.from_numpy(jax_prng.uniform(w_key, (n, m), -v, v)), torch.from_numpy(jax_prng.uniform(b_key, (n,), -v, v)))
        rng = _rng(key)
        v = np.sqrt(1 / m)
        w = torch.tensor(rng.uniform(-v, v, size=(n, m)), dtype=torch.float32)
        b = torch.tensor(rng.uniform(-v, v, size=(n,)), dtype=torch.float32)
        return w, b

    @staticmethod
    def init_params_batched(key, m, layer_sizes):
        """m independent networks at once: leaves carry the leading subdomain axis (vmap at fbpinns/trainers.py:603-607).
        With a jax-style ROOT key the reference's key derivation is followed: `key, *subkeys = split(key, m + 1)`, then
        `init_params(subkeys[i], layer_sizes)` per subdomain."""
        if jax_prng.is_key(key):
            subkeys = jax_prng.split(key, m + 1)[1:]                                   # (m, 2)
            layer_keys = jax_prng.split_batched(subkeys, len(layer_sizes) - 1)         # (m, L, 2)
            layers = []
            for l, (fi, fo) in enumerate(zip(layer_sizes[:-1], layer_sizes[1:])):
                wb = jax_prng.split_batched(layer_keys[:, l], 2)                       # (m, 2, 2): w_key, b_key
                v = np.float32(np.sqrt(np.float32(1 / fi)))
                layers.append((torch.from_numpy(jax_prng.uniform_batched(wb[:, 0], (fo, fi), -v, v)),
                               torch.from_numpy(jax_prng.uniform_batched(wb[:, 1], (fo,), -v, v))))
            return {}, {"layers": layers}
        rng = _rng(key)
        layers = []
        for fi, fo in zip(layer_sizes[:-1], layer_sizes[1:]):
            v = np.sqrt(1 / fi)
            layers.append((torch.tensor(rng.uniform(-v, v, size=(m, fo, fi)), dtype=torch.float32),
                           torch.tensor(rng.uniform(-v, v, size=(m, fo)), dtype=torch.float32)))
        return {}, {"layers": layers}

    @staticmethod
    def network_fn(params, x):
        "Host mirror for a SINGLE point (xd,) and a single subdomain's parameters (fbpinns/networks.py:61-68)"
        layers = params["trainable"]["network"]["subdomain"]["layers"]
        for w, b in layers[:-1]:
            x = torch.tanh(w @ x + b)
        w, b = layers[-1]
        return w @ x + b


def norm(mu, sd, x):
    return (x - mu) / sd


def unnorm(mu, sd, x):
    return x * sd + mu
