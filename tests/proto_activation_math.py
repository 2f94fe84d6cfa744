"""
numpy float64 transcription of the activation-jet formulas of the generic kernels' activation variants
(csrc/fbp_generic_act.cu): the reference's other Network plug-ins differ from FCN only in the activation applied to
a pre-activation jet (fbpinns/networks.py:70-194)

    kind 0  tanh(a)                    FCN, FourierFCN hidden layers      :61-68, 186-194
    kind 1  alpha * tanh(a / alpha)    AdaptiveFCN (alpha per unit)       :93-101
    kind 2  sin(a)                     SIREN, Fourier-feature layer       :125-133, 183-185
    kind 3  c * sin(o * a)             AdaptiveSIREN (c, o per unit)      :158-166

Forward, with f0..f3 = f, f', f'', f''' at a_0:      h_0 = f0,  h_k = f1 a_k,  h_kl = f1 a_kl + f2 a_k a_l
Reverse, given hbar_c:                               abar_kl = f1 hbar_kl
    abar_k  = f1 hbar_k + f2 sum_{(k,l)} (1 + delta_kl) hbar_kl a_l
    abar_0  = f1 hbar_0 + f2 sum_k hbar_k a_k + sum_kl hbar_kl (f2 a_kl + f3 a_k a_l)
    pbar    = sum_c hbar_c dh_c/dp   for the unit's own activation parameters p (alpha; c, o)
Test helper only (tests/test_oracle_and_math.py checks it against torch autograd).
"""
import numpy as np

N_EXTRA = {0: 0, 1: 1, 2: 0, 3: 2}          # trainable activation parameters per unit


def act_derivs(kind, a0, p):
    """f0, f1, f2, f3 at a0 and, per activation parameter q, the derivatives (df0/dq, df1/dq, df2/dq).
    p: tuple of arrays broadcastable to a0."""
    if kind == 0:
        t = np.tanh(a0)
        g = 1 - t * t
        return (t, g, -2 * t * g, -2 * g * (g - 2 * t * t)), ()
    if kind == 1:
        al, = p
        y = a0 / al
        t = np.tanh(y)
        g = 1 - t * t
        gp = -2 * t * g                     # dg/dy
        gpp = -2 * g * (g - 2 * t * t)      # d2g/dy2
        f = (al * t, g, gp / al, gpp / (al * al))
        dy = -y / al                        # dy/dalpha
        d0 = t + al * g * dy
        d1 = gp * dy
        d2 = gpp * dy / al - gp / (al * al)
        return f, ((d0, d1, d2),)
    if kind == 2:
        s, c = np.sin(a0), np.cos(a0)
        return (s, c, -s, -c), ()
    if kind == 3:
        cc, o = p
        s, c = np.sin(o * a0), np.cos(o * a0)
        f = (cc * s, cc * o * c, -cc * o * o * s, -cc * o ** 3 * c)
        dc = (s, o * c, -o * o * s)
        do = (cc * a0 * c, cc * (c - o * a0 * s), -cc * (2 * o * s + o * o * a0 * c))
        return f, (dc, do)
    raise ValueError(kind)


def act_forward(kind, a, p, order, i1, i2):
    "a: (..., C) pre-activation jets -> h (..., C)"
    (f0, f1, f2, _), _ = act_derivs(kind, a[..., 0], p)
    h = np.empty_like(a)
    h[..., 0] = f0
    for c in range(1, a.shape[-1]):
        if order[c] == 1:
            h[..., c] = f1 * a[..., c]
        else:
            h[..., c] = f1 * a[..., c] + f2 * a[..., i1[c]] * a[..., i2[c]]
    return h


def act_backward(kind, a, p, hbar, order, i1, i2):
    "-> abar (..., C), [pbar per activation parameter (...)]"
    (f0, f1, f2, f3), dps = act_derivs(kind, a[..., 0], p)
    C = a.shape[-1]
    ab = f1[..., None] * hbar
    ab0 = f1 * hbar[..., 0]
    for c in range(1, C):
        if order[c] == 1:
            ab0 = ab0 + f2 * hbar[..., c] * a[..., c]
        else:
            k, l = i1[c], i2[c]
            ab0 = ab0 + hbar[..., c] * (f2 * a[..., c] + f3 * a[..., k] * a[..., l])
            ab[..., k] = ab[..., k] + f2 * hbar[..., c] * a[..., l]
            ab[..., l] = ab[..., l] + f2 * hbar[..., c] * a[..., k]
    ab[..., 0] = ab0
    pbar = []
    for (d0, d1, d2) in dps:
        v = hbar[..., 0] * d0
        for c in range(1, C):
            if order[c] == 1:
                v = v + hbar[..., c] * d1 * a[..., c]
            else:
                v = v + hbar[..., c] * (d1 * a[..., c] + d2 * a[..., i1[c]] * a[..., i2[c]])
        pbar.append(v)
    return ab, pbar
