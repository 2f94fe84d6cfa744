"""
Network plug-ins (API of fbpinns/networks.py:14-194, 197-201): FCN (tanh; tiled FFMA2 and tensor kernel families) and the
reference's other networks AdaptiveFCN, SIREN, AdaptiveSIREN, FourierFCN (activation variants of the generic kernel family,
csrc/fbp_generic_act.cu).  `kernel_layers` is what the kernels see of a network.
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
    ACTIVATION = "tanh"

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


def _layers_from_subkeys(subkeys, layer_sizes, gain, n_extra):
    "row i: the layers `init_params(subkeys[i], layer_sizes)` of the reference draws (split per layer, then w_key / b_key)"
    layers = []
    layer_keys = jax_prng.split_batched(subkeys, len(layer_sizes) - 1)
    for l, (fi, fo) in enumerate(zip(layer_sizes[:-1], layer_sizes[1:])):
        wb = jax_prng.split_batched(layer_keys[:, l], 2)
        v = np.float32(np.sqrt(np.float32(gain / fi)))
        w = torch.from_numpy(jax_prng.uniform_batched(wb[:, 0], (fo, fi), -v, v))
        b = torch.from_numpy(jax_prng.uniform_batched(wb[:, 1], (fo,), -v, v))
        layers.append((w, b) + tuple(torch.ones_like(b) for _ in range(n_extra)))
    return layers


def _layers_batched(key, m, layer_sizes, gain, n_extra):
    """m independent stacks of (W (m,out,in), b (m,out), n_extra x ones (m,out)) with W, b ~ U(-v, v), v = sqrt(gain/in):
    the draws of every reference network's `_random_layer_params`, for a numpy Generator / seed or a threefry ROOT key
    (then `key, *subkeys = split(key, m+1)` as at fbpinns/trainers.py:603)."""
    if jax_prng.is_key(key):
        return _layers_from_subkeys(jax_prng.split(key, m + 1)[1:], layer_sizes, gain, n_extra)
    rng = _rng(key)
    layers = []
    for fi, fo in zip(layer_sizes[:-1], layer_sizes[1:]):
        v = np.sqrt(gain / fi)
        w = torch.tensor(rng.uniform(-v, v, size=(m, fo, fi)), dtype=torch.float32)
        b = torch.tensor(rng.uniform(-v, v, size=(m, fo)), dtype=torch.float32)
        layers.append((w, b) + tuple(torch.ones_like(b) for _ in range(n_extra)))
    return layers


def _single(layers_batched):
    return [tuple(t[0] for t in leaf) for leaf in layers_batched]


class AdaptiveFCN(Network):
    "Fully connected network with adaptive activations a * tanh(x / a), a per unit (fbpinns/networks.py:70-101)"
    ACTIVATION, N_EXTRA, GAIN = "adaptive_tanh", 1, 1.0

    @classmethod
    def init_params(cls, key, layer_sizes):
        if jax_prng.is_key(key):          # a single network draws from its own key: no subdomain split
            return {}, {"layers": _single(_layers_from_subkeys(key[None], layer_sizes, cls.GAIN, cls.N_EXTRA))}
        return {}, {"layers": _single(_layers_batched(key, 1, layer_sizes, cls.GAIN, cls.N_EXTRA))}

    @classmethod
    def init_params_batched(cls, key, m, layer_sizes):
        return {}, {"layers": _layers_batched(key, m, layer_sizes, cls.GAIN, cls.N_EXTRA)}

    @staticmethod
    def network_fn(params, x):
        layers = params["trainable"]["network"]["subdomain"]["layers"]
        for w, b, a in layers[:-1]:
            x = a * torch.tanh((w @ x + b) / a)
        w, b, _ = layers[-1]
        return w @ x + b


class SIREN(AdaptiveFCN):
    "Fully connected network with sin activations, U(+-sqrt(6/in)) initialisation (fbpinns/networks.py:103-133)"
    ACTIVATION, N_EXTRA, GAIN = "sin", 0, 6.0

    @staticmethod
    def network_fn(params, x):
        layers = params["trainable"]["network"]["subdomain"]["layers"]
        for w, b in layers[:-1]:
            x = torch.sin(w @ x + b)
        w, b = layers[-1]
        return w @ x + b


class AdaptiveSIREN(AdaptiveFCN):
    "Fully connected network with adaptive sin activations c * sin(o * x), c and o per unit (fbpinns/networks.py:135-166)"
    ACTIVATION, N_EXTRA, GAIN = "adaptive_sin", 2, 6.0

    @staticmethod
    def network_fn(params, x):
        layers = params["trainable"]["network"]["subdomain"]["layers"]
        for w, b, c, o in layers[:-1]:
            x = c * torch.sin(o * (w @ x + b))
        w, b, _, _ = layers[-1]
        return w @ x + b


class FourierFCN(FCN):
    """FCN on Fourier features [sin(omega x), cos(omega x)] with a static omega per subdomain,
    omega = 2 pi (mu + sd N(0,1)) of shape (n_features, xd)  (fbpinns/networks.py:168-194)"""
    ACTIVATION = "fourier_tanh"

    @staticmethod
    def _omega(key, m, xd, mu, sd, n_features):
        if jax_prng.is_key(key):
            raise NotImplementedError("threefry initialisation of FourierFCN (needs XLA's float32 erf_inv bits)")
        rng = _rng(key)
        return torch.tensor(2 * np.pi * (mu + sd * rng.standard_normal((m, n_features, xd))), dtype=torch.float32)

    @staticmethod
    def init_params(key, layer_sizes, mu, sd, n_features):
        st, tr = FourierFCN.init_params_batched(key, 1, layer_sizes, mu, sd, n_features)
        return {"omega": st["omega"][0]}, {"layers": _single(tr["layers"])}

    @staticmethod
    def init_params_batched(key, m, layer_sizes, mu, sd, n_features):
        rng = _rng(key)
        omega = FourierFCN._omega(rng, m, layer_sizes[0], mu, sd, n_features)
        sizes = [2 * n_features] + list(layer_sizes)[1:]
        return {"omega": omega}, {"layers": _layers_batched(rng, m, sizes, 1.0, 0)}

    @staticmethod
    def network_fn(params, x):
        omega = params["static"]["network"]["subdomain"]["omega"]
        layers = params["trainable"]["network"]["subdomain"]["layers"]
        x = omega @ x
        x = torch.cat([torch.sin(x), torch.cos(x)])
        for w, b in layers[:-1]:
            x = torch.tanh(w @ x + b)
        w, b = layers[-1]
        return w @ x + b


def kernel_layers(network, all_params, device=None):
    """What the kernels need of a network: (activation tag of engine.Plan, layer sizes, [(W, b, activation parameters...)])
    with every leaf carrying the leading subdomain axis.  For FourierFCN the static feature layer becomes layer 0,
    W0 = [omega; omega], b0 = [0; pi/2], so that sin(W0 z + b0) = [sin(omega z), cos(omega z)] (fbpinns/networks.py:183-185)."""
    act = getattr(network, "ACTIVATION", None)
    if act is None:
        raise NotImplementedError(f"{network} is not implemented by the B200 kernels")
    # leaves may be torch tensors or numpy arrays (a checkpoint read by util.checkpoint.load_model holds numpy leaves)
    as_t = lambda t: t if torch.is_tensor(t) else torch.as_tensor(np.asarray(t))
    to = (lambda t: as_t(t).to(device, torch.float32)) if device is not None else (lambda t: as_t(t).float())
    layers = [tuple(to(t) for t in leaf) for leaf in all_params["trainable"]["network"]["subdomain"]["layers"]]
    if act == "fourier_tanh":
        omega = to(all_params["static"]["network"]["subdomain"]["omega"])           # (m, n_features, xd)
        m, nf, _ = omega.shape
        b0 = torch.cat([torch.zeros((m, nf), dtype=omega.dtype, device=omega.device),
                        torch.full((m, nf), float(np.pi / 2), dtype=omega.dtype, device=omega.device)], dim=1)
        layers = [(torch.cat([omega, omega], dim=1).contiguous(), b0)] + layers
    sizes = [int(layers[0][0].shape[2])] + [int(leaf[0].shape[1]) for leaf in layers]
    return act, sizes, layers


def from_kernel_layers(network, layers):
    "inverse of kernel_layers for the trainable leaves (drops FourierFCN's static feature layer)"
    return layers[1:] if getattr(network, "ACTIVATION", None) == "fourier_tanh" else layers


def norm(mu, sd, x):
    return (x - mu) / sd


def unnorm(mu, sd, x):
    return x * sd + mu
