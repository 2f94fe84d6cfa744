"""
numpy float64 transcription of the formulas the CUDA kernels implement (forward jets, Leibniz product with the
window jets, row sums + quotient rule, and the hand-derived reverse pass).  Test helper only: it lets the CPU
suite check the kernel MATHS against the oracle's autograd before any GPU time is spent.
"""
import numpy as np


def comp_tables(jet):
    C = jet.C
    order = [len(p) for p in jet.comps]
    i1, i2 = [0] * C, [0] * C
    for c, p in enumerate(jet.comps):
        if len(p) == 2:
            i1[c], i2[c] = jet.index[(p[0],)], jet.index[(p[1],)]
    return order, i1, i2


def window_jets(jet, z, isd, flag):
    "z, isd: (s, xd); flag: (s,) -> (s, C)"
    s, xd = z.shape
    c_, s_ = np.cos(np.pi * z), np.sin(np.pi * z)
    A = 0.5 * (1 + c_)
    kap = np.pi * isd
    f, f1, f2 = A * A, -kap * s_ * A, -kap * kap * (2 * c_ - 1) * A
    w = np.zeros((s, jet.C))
    for c, p in enumerate(jet.comps):
        prod = flag.copy()
        for d in range(xd):
            cnt = sum(1 for a in p if a == d)
            prod = prod * (f[:, d] if cnt == 0 else f1[:, d] if cnt == 1 else f2[:, d])
        w[:, c] = prod + (1 - flag) if c == 0 else prod
    return w


def pair_setup(x_take, ss):
    "ss: (s, 2*xd+3) static record per pair"
    xd = x_take.shape[1]
    lo, hi = ss[:, :xd], ss[:, xd:2 * xd]
    mu, sd = (hi + lo) * 0.5, (hi - lo) * 0.5
    isd = 1.0 / sd
    return (x_take - mu) * isd, isd, ss[:, 2 * xd], ss[:, 2 * xd + 1], ss[:, 2 * xd + 2]


def pair_forward(jet, x_take, ss, layers_take):
    """layers_take: [(W (s,out,in), b (s,out))]. Returns N (s, C*ud), cache."""
    order, i1, i2 = comp_tables(jet)
    C, ud = jet.C, jet.ud
    z, isd, flag, un_mu, un_sd = pair_setup(x_take, ss)
    s, xd = z.shape
    hin = np.zeros((s, C, xd))
    hin[:, 0, :] = z
    for c, p in enumerate(jet.comps):
        if len(p) == 1:
            hin[:, c, p[0]] = isd[:, p[0]]
    hs = [hin]
    L = len(layers_take)
    for l, (W, b) in enumerate(layers_take):
        a = np.einsum("soi,sci->sco", W, hs[-1])
        a[:, 0, :] += b
        if l == L - 1:
            r = a
            break
        t = np.tanh(a[:, 0, :])
        g = 1 - t * t
        h = np.zeros_like(a)
        h[:, 0] = t
        for c in range(1, C):
            if order[c] == 1:
                h[:, c] = g * a[:, c]
            else:
                h[:, c] = g * (a[:, c] - 2 * t * a[:, i1[c]] * a[:, i2[c]])
        hs.append(h)
    u = un_sd[:, None, None] * r
    u[:, 0, :] += un_mu[:, None]
    w = window_jets(jet, z, isd, flag)
    N = np.zeros((s, C, ud))
    for c in range(C):
        if order[c] == 0:
            N[:, c] = u[:, 0] * w[:, 0:1]
        elif order[c] == 1:
            N[:, c] = u[:, c] * w[:, 0:1] + u[:, 0] * w[:, c:c + 1]
        else:
            N[:, c] = (u[:, c] * w[:, 0:1] + u[:, i1[c]] * w[:, i2[c]:i2[c] + 1] + u[:, i2[c]] * w[:, i1[c]:i1[c] + 1]
                       + u[:, 0] * w[:, c:c + 1])
    return N.reshape(s, C * ud), dict(hs=hs, w=w, un_sd=un_sd, z=z, isd=isd)


def reduce_forward(jet, N_ref, dsum, takes, n):
    "N_ref: (s, C*ud) in reference pair order"
    order, i1, i2 = comp_tables(jet)
    m_take, n_take, p_take, np_take, npou = takes
    C, ud = jet.C, jet.ud
    q = len(np_take)
    Nr = np.zeros((q, C, ud))
    np.add.at(Nr, p_take, N_ref.reshape(-1, C, ud))
    D = dsum
    invD = 1.0 / D[:, 0]
    u = np.zeros((q, C, ud))
    u[:, 0] = Nr[:, 0] * invD[:, None]
    for c in range(1, C):
        if order[c] == 1:
            u[:, c] = (Nr[:, c] - u[:, 0] * D[:, c:c + 1]) * invD[:, None]
    for c in range(1, C):
        if order[c] == 2:
            u[:, c] = (Nr[:, c] - u[:, i1[c]] * D[:, i2[c]:i2[c] + 1] - u[:, i2[c]] * D[:, i1[c]:i1[c] + 1]
                       - u[:, 0] * D[:, c:c + 1]) * invD[:, None]
    out = np.zeros((n, C, ud))
    np.add.at(out, np_take, u)
    return (out / npou).reshape(n, C * ud)


def reduce_backward(jet, ubar, dsum, takes):
    order, i1, i2 = comp_tables(jet)
    m_take, n_take, p_take, np_take, npou = takes
    C, ud = jet.C, jet.ud
    q = len(np_take)
    ub = ubar.reshape(-1, C, ud)[np_take] / npou
    ub = ub.copy()
    D = dsum
    invD = (1.0 / D[:, 0])[:, None]
    nb = np.zeros((q, C, ud))
    for c in range(1, C):
        if order[c] == 2:
            t = ub[:, c] * invD
            nb[:, c] = t
            ub[:, i1[c]] -= t * D[:, i2[c]:i2[c] + 1]
            ub[:, i2[c]] -= t * D[:, i1[c]:i1[c] + 1]
            ub[:, 0] -= t * D[:, c:c + 1]
    for c in range(1, C):
        if order[c] == 1:
            t = ub[:, c] * invD
            nb[:, c] = t
            ub[:, 0] -= t * D[:, c:c + 1]
    nb[:, 0] = ub[:, 0] * invD
    return nb.reshape(q, C * ud)


def pair_backward(jet, layers_take, cache, G, m_take, m_active):
    """G: (s, C*ud) cotangent of N per pair. Returns grads [(gW (m_active,out,in), gb (m_active,out))]."""
    order, i1, i2 = comp_tables(jet)
    C, ud = jet.C, jet.ud
    hs, w, un_sd = cache["hs"], cache["w"], cache["un_sd"]
    s = G.shape[0]
    G = G.reshape(s, C, ud)
    ub = np.zeros((s, C, ud))
    for c in range(C):
        ub[:, 0] += G[:, c] * w[:, c:c + 1]
        if order[c] == 1:
            ub[:, c] += G[:, c] * w[:, 0:1]
        elif order[c] == 2:
            ub[:, c] += G[:, c] * w[:, 0:1]
            ub[:, i1[c]] += G[:, c] * w[:, i2[c]:i2[c] + 1]
            ub[:, i2[c]] += G[:, c] * w[:, i1[c]:i1[c] + 1]
    abar = un_sd[:, None, None] * ub                      # cotangent of the output layer pre-activation jets
    grads = []
    L = len(layers_take)
    for l in range(L - 1, -1, -1):
        W, b = layers_take[l]
        hin = hs[l]
        gW_pair = np.einsum("sco,sci->soi", abar, hin)
        gb_pair = abar[:, 0, :]
        gW = np.zeros((m_active,) + W.shape[1:])
        gb = np.zeros((m_active,) + b.shape[1:])
        act = m_take < m_active
        np.add.at(gW, m_take[act], gW_pair[act])
        np.add.at(gb, m_take[act], gb_pair[act])
        grads.append((gW, gb))
        if l > 0:
            hbar = np.einsum("soi,sco->sci", W, abar)
            t = hin[:, 0]
            g = 1 - t * t
            ab = g[:, None, :] * hbar
            ab0 = g * hbar[:, 0]
            for c in range(1, C):
                if order[c] == 1:
                    ab0 = ab0 - 2 * t * hbar[:, c] * hin[:, c]
                else:
                    a1, a2 = i1[c], i2[c]
                    ab0 = ab0 - 2 * hbar[:, c] * (t * hin[:, c] + hin[:, a1] * hin[:, a2])
                    ab[:, a1] -= 2 * t * hbar[:, c] * hin[:, a2]
                    ab[:, a2] -= 2 * t * hbar[:, c] * hin[:, a1]
            ab[:, 0] = ab0
            abar = ab
    return grads[::-1]
