"""
Host mirror of the window function the kernels implement (API parity with fbpinns/windows.py:25-35).
The device implementation (value and x-derivatives) is fbp_window_dim in csrc/fbp_common.cuh.
"""
import math

import torch


def cosine(xmin, xmax, x):
    "window function, for a SINGLE point with shape (xd,) — or batched over a leading axis"
    mu, sd = (xmin + xmax) / 2, (xmax - xmin) / 2
    ws = ((1 + torch.cos(math.pi * (x - mu) / sd)) / 2) ** 2
    ws = (x - xmin >= 0).to(x.dtype) * (xmax - x >= 0).to(x.dtype) * ws
    return torch.prod(ws, dim=-1, keepdim=True)
