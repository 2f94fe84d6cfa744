"""
Jet specification: which derivative components of the subdomain sum the kernels must produce for one
constraint, derived from the reference's `required_ujs` / `get_jmaps` trie (fbpinns/trainers.py:63-107), and
the host-side application of `Problem.constraining_fn` to those jets.

The reference differentiates THROUGH the constraining operator with nested forward-mode jvps
(fbpinns/trainers.py:174, 213-247).  Here the kernels return the jets of the unconstrained sum; the constrained
ujs are recovered exactly by running the same nested-jvp chain (`get_ujs`, same structure as the reference's
`_get_ujs`) on  x' -> constraining_fn(all_params, x', T(x'))  at x' = x, where T is the per-point Taylor
polynomial of the unconstrained solution built from the jets.  The chain rule only ever touches derivatives of
T up to the requested order along the requested axes, which are exactly the jet components.
"""

import torch
from torch.func import jvp


def get_jmaps(required_ujs):
    """Same contract as the reference's get_jmaps (fbpinns/trainers.py:63-107): returns (nodes, leaves, jac_is).
    nodes[i] = ((parent function index, axis), path, is_leaf); function 0 is u itself."""
    root = {}
    for _, path in required_ujs:
        t = root
        for ix in path:
            t = t.setdefault(ix, {})
    nodes = []

    def visit(t, parent, prefix):
        for ix in t:
            p = prefix + (ix,)
            nodes.append(((parent, ix), p, int(not t[ix])))
            if t[ix]:
                visit(t[ix], len(nodes), p)
    visit(root, 0, ())
    nodes = tuple(nodes)
    leaves = tuple((i + 1, n[1]) for i, n in enumerate(nodes) if n[2]) or ((0, ()),)
    jac_is = tuple(([lf[1][:len(p)] for lf in leaves].index(tuple(p)), len(p), iu) for iu, p in required_ujs)
    return nodes, leaves, jac_is


class JetSpec:
    """Closed, canonically ordered set of jet components for one constraint.

    comps[c] is a sorted axis tuple: () value, (k,) first order, (k, l) with k <= l second order.
    Order: value, first-order components by axis, second-order components lexicographically."""

    def __init__(self, required_ujs, xd, ud):
        self.required_ujs = tuple((int(iu), tuple(int(i) for i in p)) for iu, p in required_ujs)
        self.xd, self.ud = int(xd), int(ud)
        want = {()}
        for iu, path in self.required_ujs:
            if not 0 <= iu < ud:
                raise ValueError(f"required_ujs: solution index {iu} out of range for ud={ud}")
            if any(not 0 <= i < xd for i in path):
                raise ValueError(f"required_ujs: axis out of range in {path} for xd={xd}")
            if len(path) > 2:
                raise NotImplementedError(
                    f"required_ujs asks for a derivative of order {len(path)} ({path}); the B200 kernels carry jets "
                    f"up to order 2 (every reference problem needs <= 2, fbpinns/problems.py)")
            sp = tuple(sorted(path))
            want.add(sp)
            for i in sp:            # closure: every order-2 component needs its order-1 components
                want.add((i,))
        self.comps = sorted(want, key=lambda p: (len(p), p))
        self.index = {p: c for c, p in enumerate(self.comps)}
        self.C = len(self.comps)
        self.jmaps = get_jmaps(self.required_ujs)

    @property
    def comp_k(self):
        return [p[0] if len(p) >= 1 else -1 for p in self.comps]

    @property
    def comp_l(self):
        return [p[1] if len(p) == 2 else -1 for p in self.comps]

    def key(self):
        return (self.xd, self.ud, tuple(self.comps))

    def column(self, iu, path):
        """Column of `ujets` (n, C*ud) holding d^path u_iu."""
        return self.index[tuple(sorted(path))] * self.ud + iu

    # ---- ujs without / with a constraining operator -------------------------------------------------------
    def ujs_plain(self, ujets):
        return [ujets[:, self.column(iu, p):self.column(iu, p) + 1] for iu, p in self.required_ujs]

    def taylor(self, ujets, x0, x):
        """T(x) = sum_c jet_c (x-x0)^alpha_c / alpha_c!  -> (n, ud)."""
        n = ujets.shape[0]
        J = ujets.view(n, self.C, self.ud)
        dx = x - x0
        out = J[:, 0, :]
        for c, p in enumerate(self.comps):
            if len(p) == 1:
                out = out + J[:, c, :] * dx[:, p[0]:p[0] + 1]
            elif len(p) == 2:
                k, l = p
                coef = 0.5 if k == l else 1.0
                out = out + coef * J[:, c, :] * dx[:, k:k + 1] * dx[:, l:l + 1]
        return out

    def ujs_constrained(self, ujets, x_batch, constraining_fn, all_params):
        """ujs of constraining_fn(all_params, x, u(x)) by the reference's nested-jvp chain on the local Taylor
        model (exact: see module docstring)."""
        x0 = x_batch.detach()

        def u_fn(x):
            return constraining_fn(all_params, x, self.taylor(ujets, x0, x)), ()
        return get_ujs(x0, self.jmaps, u_fn)


class AffineConstraining:
    """Fast path for constraining operators that are affine in u with point-wise coefficients,
           constraining_fn(all_params, x, u) = A(x) * u + B(x)        (ud = 1)
    which covers every hard-boundary-condition ansatz of the reference problems (fbpinns/problems.py:193-199,
    312-318, 381-398) and Poisson2D.  The jets of A and B are static while the collocation points do not change, so
    they are computed ONCE per active-set change with the generic nested-jvp machinery (`JetSpec.ujs_constrained`
    on the probes u = 0 and u = 1) and the constrained ujs follow from the Leibniz rule on the jets of u —
    a dozen elementwise operations per step instead of several hundred.

    `build` returns None (caller falls back to the generic path) unless the operator passes a numerical affinity
    check on random jets at every collocation point."""

    def __init__(self, jet, Aj, Bj):
        self.jet, self.Aj, self.Bj = jet, Aj, Bj          # (n, C) jets of A and B in the jet's component order

    @staticmethod
    def build(jet, x_batch, constraining_fn, all_params, rtol=None):
        if jet.ud != 1 or x_batch.shape[0] == 0:
            return None
        if rtol is None:            # a non-affine operator deviates by O(1); the bound only has to clear round-off
            rtol = 1e-9 if x_batch.dtype == torch.float64 else 1e-4
        n, C = x_batch.shape[0], jet.C
        full = JetSpec(tuple((0, p) for p in jet.comps), jet.xd, 1)        # every component, in the same order
        assert full.comps == jet.comps

        def jets_of(ujets):
            with torch.no_grad():
                cols = full.ujs_constrained(ujets, x_batch, constraining_fn, all_params)
            return torch.cat(cols, dim=1)
        zeros = torch.zeros((n, C), dtype=x_batch.dtype, device=x_batch.device)
        ones = zeros.clone()
        ones[:, 0] = 1.0
        try:
            Bj = jets_of(zeros)
            Aj = jets_of(ones) - Bj
            gen = torch.Generator(device="cpu").manual_seed(1234)
            probe = torch.randn((n, C), generator=gen, dtype=x_batch.dtype).to(x_batch.device)
            want = jets_of(probe)
        except Exception:
            return None
        aff = AffineConstraining(jet, Aj, Bj)
        got = torch.cat(aff._apply(probe, full.required_ujs), dim=1)
        scale = want.abs().amax(dim=0).clamp_min(1e-30)
        if not torch.isfinite(want).all() or ((got - want).abs().amax(dim=0) / scale).max().item() > rtol:
            return None
        return aff

    def _apply(self, ujets, required):
        jet, A, B = self.jet, self.Aj, self.Bj
        col = lambda T, p: T[:, jet.index[p]:jet.index[p] + 1]
        out = []
        for _, path in required:
            p = tuple(sorted(path))
            if len(p) == 0:
                v = col(A, ()) * col(ujets, ()) + col(B, ())
            elif len(p) == 1:
                v = col(A, p) * col(ujets, ()) + col(A, ()) * col(ujets, p) + col(B, p)
            else:
                k, l = (p[0],), (p[1],)
                v = (col(A, p) * col(ujets, ()) + col(A, k) * col(ujets, l) + col(A, l) * col(ujets, k)
                     + col(A, ()) * col(ujets, p) + col(B, p))
            out.append(v)
        return out

    def ujs(self, ujets):
        "constrained ujs in the order of the constraint's required_ujs"
        return self._apply(ujets, self.jet.required_ujs)


def _jacfwd(f, v):
    def jacfun(x):
        y, j, aux = jvp(f, (x,), (v,), has_aux=True)
        return j, aux + (y,)
    return jacfun


def get_ujs(x_batch, jmaps, u_fn):
    """Chained forward-mode derivatives with one-hot tangents per point; same contract as the reference's
    _get_ujs (fbpinns/trainers.py:213-239)."""
    nodes, leaves, jac_is = jmaps
    n, xd = x_batch.shape
    eye = torch.eye(xd, dtype=x_batch.dtype, device=x_batch.device)
    fs = [u_fn]
    for (ni, ix), _, _ in nodes:
        fs.append(_jacfwd(fs[ni], eye[ix].expand(n, xd)))
    jacs = []
    for ie, _ in leaves:
        fin, jac = fs[ie](x_batch)
        jacs.append(jac + (fin,))
    return [jacs[il][io][:, iu:iu + 1] for il, io, iu in jac_is]
