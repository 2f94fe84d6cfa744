"""Case definitions shared by make_golden_extra.py (reference side) and tests/test_golden_extra.py (our side)."""
import numpy as np


def _ws(subdomain_xs, width):          # fbpinns/constants.py:21-22
    return [width * np.min(np.diff(x)) * np.ones_like(x) for x in subdomain_xs]


def multilevel_setup():
    "two levels on [0,1]^2: level 0 = one window-less subdomain covering the domain, level 1 = 3 x 2 overlapping boxes"
    xs0 = [np.array([0.5]), np.array([0.5])]
    ws0 = [np.array([1.4]), np.array([1.4])]
    xs1 = [np.linspace(0, 1, 3), np.linspace(0, 1, 2)]
    ws1 = _ws(xs1, 2.9)
    g = [np.linspace(0, 1, 11), np.linspace(0, 1, 9)]
    return dict(dkw=dict(subdomain_xss=[xs0, xs1], subdomain_wss=[ws0, ws1], unnorm=(0.2, 1.3)),
                x=np.stack(np.meshgrid(*g, indexing="ij"), -1).reshape(-1, 2).astype(np.float32),
                layer_sizes=[2, 8, 8, 1], req=((0, ()), (0, (0,)), (0, (1,)), (0, (0, 0)), (0, (1, 1))))


def multilevel_masks(m):
    a = np.ones(m, dtype=int)
    b = np.array([1, 1, 2, 0, 1, 2, 1][:m], dtype=int)
    return [a, b]


def ho1d_setup():
    xs = [np.linspace(0, 1, 7)]
    return dict(dkw=dict(subdomain_xs=xs, subdomain_ws=_ws(xs, 2.5), unnorm=(0., 1.)), pkw=dict(d=2, w0=20),
                x_phys=np.linspace(0, 1, 31).reshape(-1, 1).astype(np.float32), layer_sizes=[1, 8, 1])


def ho1d_masks(m):
    a = np.ones(m, dtype=int)
    b = np.array([2, 2, 1, 1, 0, 0, 0][:m], dtype=int)       # boundary point x = 0 lies only in FIXED subdomains
    c = np.array([0, 0, 0, 0, 1, 1, 2][:m], dtype=int)       # ... and here in none of the active ones: empty constraint
    return [a, b, c]
