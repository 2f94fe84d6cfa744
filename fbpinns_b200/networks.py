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


class FCN(Network):
    "Fully connected network"

    @staticmethod
    def init_params(key, layer_sizes):
        """`key` is a numpy Generator or an integer seed (jax.random keys do not exist here; parameter
        initialisation parity with the reference is version-dependent and unpinned)."""
        rng = _rng(key)
        params = [FCN._random_layer_params(rng, m, n) for m, n in zip(layer_sizes[:-1], layer_sizes[1:])]
        return {}, {"layers": params}

    @staticmethod
    def _random_layer_params(key, m, n):
        "U(-1/sqrt(fan_in), 1/sqrt(fan_in)) weights (n, m) and biases (n,) — fbpinns/networks.py:48-58"
        rng = _rng(key)
        v = np.sqrt(1 / m)
        w = torch.tensor(rng.uniform(-v, v, size=(n, m)), dtype=torch.float32)
        b = torch.tensor(rng.uniform(-v, v, size=(n,)), dtype=torch.float32)
        return w, b

    @staticmethod
    def init_params_batched(key, m, layer_sizes):
        "m independent networks at once: leaves carry the leading subdomain axis (vmap at fbpinns/trainers.py:603-607)"
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
