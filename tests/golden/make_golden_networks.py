"""
Golden fixture for the reference's Network plug-ins (fbpinns/networks.py:36-194): FCN, AdaptiveFCN, SIREN,
AdaptiveSIREN, FourierFCN `network_fn` EXECUTED FROM THE REFERENCE'S OWN SOURCE under the numpy-backed jax shim of
make_golden_shim.py, in float64, on explicit random parameters (non-trivial activation parameters so that every
term is exercised).  Output: tests/golden/refnetworks.npz (inputs and outputs).

Run:  python tests/golden/make_golden_networks.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_shim import install_shim, AtArray          # noqa: E402

LAYER_SIZES = [2, 6, 5, 1]
N_FEATURES = 3


def main():
    ref = install_shim()
    N = ref.networks
    rng = np.random.default_rng(7)
    A = lambda a: np.asarray(a, dtype=np.float64).view(AtArray)
    x = rng.uniform(-1, 1, (9, LAYER_SIZES[0]))
    out = {"x": x, "layer_sizes": np.array(LAYER_SIZES), "n_features": np.array(N_FEATURES)}

    def layers(sizes, n_extra):
        ls = []
        for fi, fo in zip(sizes[:-1], sizes[1:]):
            v = np.sqrt(1 / fi)
            leaf = [rng.uniform(-v, v, (fo, fi)), rng.uniform(-v, v, (fo,))]
            leaf += [rng.uniform(0.6, 1.4, (fo,)) for _ in range(n_extra)]
            ls.append(tuple(leaf))
        return ls

    for name, cls, n_extra in [("fcn", N.FCN, 0), ("adaptive_fcn", N.AdaptiveFCN, 1), ("siren", N.SIREN, 0),
                               ("adaptive_siren", N.AdaptiveSIREN, 2), ("fourier", N.FourierFCN, 0)]:
        static = {}
        sizes = list(LAYER_SIZES)
        if name == "fourier":
            omega = 2 * np.pi * (0.1 + 0.8 * rng.standard_normal((N_FEATURES, LAYER_SIZES[0])))
            static = {"network": {"subdomain": {"omega": A(omega)}}}
            sizes = [2 * N_FEATURES] + sizes[1:]
            out[f"{name}_omega"] = omega
        ls = layers(sizes, n_extra)
        params = {"static": static, "trainable": {"network": {"subdomain": {"layers": [tuple(A(t) for t in leaf) for leaf in ls]}}}}
        y = np.stack([np.asarray(cls.network_fn(params, A(xi))) for xi in x])
        out[f"{name}_y"] = y
        for l, leaf in enumerate(ls):
            for i, t in enumerate(leaf):
                out[f"{name}_l{l}_{i}"] = t
    np.savez(os.path.join(HERE, "refnetworks.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("_y")})


if __name__ == "__main__":
    main()
