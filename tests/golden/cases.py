"""Case definitions shared by make_golden_shim.py (reference side) and tests/test_golden_reference.py (our side)."""
import numpy as np

NAMES = ["ho1d_hardbc", "burgers2d", "wave3d"]


def _ws(subdomain_xs, width):          # fbpinns/constants.py:21-22
    return [width * np.min(np.diff(x)) * np.ones_like(x) for x in subdomain_xs]


def case_setup(name):
    if name == "ho1d_hardbc":
        xs = [np.linspace(0, 1, 9)]
        return dict(dkw=dict(subdomain_xs=xs, subdomain_ws=_ws(xs, 2.7), unnorm=(0.1, 1.5)),
                    problem="HarmonicOscillator1DHardBC", pkw=dict(d=2, w0=20, sd=0.1),
                    x=np.linspace(0, 1, 41).reshape(-1, 1).astype(np.float32), layer_sizes=[1, 8, 1],
                    req=((0, ()), (0, (0,)), (0, (0, 0))))
    if name == "burgers2d":
        xs = [np.linspace(-1, 1, 4), np.linspace(0, 1, 3)]
        g = [np.linspace(-1, 1, 13), np.linspace(0, 1, 11)]
        return dict(dkw=dict(subdomain_xs=xs, subdomain_ws=_ws(xs, 2.9), unnorm=(0., 3.)),
                    problem="BurgersEquation2D", pkw=dict(),
                    x=np.stack(np.meshgrid(*g, indexing="ij"), -1).reshape(-1, 2).astype(np.float32),
                    layer_sizes=[2, 8, 8, 1], req=((0, ()), (0, (0,)), (0, (1,)), (0, (0, 0))))
    xs = [np.linspace(-1, 1, 3), np.linspace(-1, 1, 2), np.linspace(0, 1, 3)]
    g = [np.linspace(-1, 1, 7), np.linspace(-1, 1, 6), np.linspace(0, 1, 5)]
    return dict(dkw=dict(subdomain_xs=xs, subdomain_ws=_ws(xs, 2.9), unnorm=(0., 1.)),
                problem="WaveEquationGaussianVelocity3D", pkw=dict(),
                x=np.stack(np.meshgrid(*g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32),
                layer_sizes=[3, 8, 1], req=((0, (0, 0)), (0, (1, 1)), (0, (2, 2))))
