"""
Active schedulers: which subdomains train at which step (public names and semantics of fbpinns/schedulers.py:16-161).

A scheduler is an iterable of length n_steps yielding, per training step, either None ("active set unchanged") or an
int array of length m with 0 = inactive (still trained if it overlaps active subdomains), 1 = active, 2 = fixed.
It is pure float64 host arithmetic (~100 lines in the reference) whose OUTPUT the trainer consumes bit-exactly, so it
is host code here too; `tests/golden/schedulers.npz` (produced by the reference's own module) pins every state.

The radial schedulers share one implementation: a front at radius r(i) = r0 + (r1 - r0) * i / n_steps sweeps outwards
from a point / line / plane; a subdomain (taken as the box between the centre lines of its overlaps, `xmins0/xmaxs0`)
switches on when the front enters the shell [nearest, farthest) of its distances to the origin set and is frozen
when the front has left it.
"""
import numpy as np


class ActiveScheduler:
    "base class: knows n_steps, the number of subdomains m and the dimensionality xd"

    def __init__(self, all_params, n_steps):
        d = all_params["static"]["decomposition"]
        self.n_steps, self.m, self.xd = n_steps, d["m"], d["xd"]

    def __len__(self):
        return self.n_steps

    def __iter__(self):
        raise NotImplementedError


class AllActiveSchedulerND(ActiveScheduler):
    "every subdomain trains from the first step on"

    def __iter__(self):
        yield np.ones(self.m, dtype=int)
        for _ in range(1, self.n_steps):
            yield None


def _shell_of_boxes(origin, lo, hi):
    """Nearest and farthest distance from `origin` (cd,) to each box [lo, hi] (m, cd).
    Same float64 operations as the reference's `_get_radii` (fbpinns/schedulers.py:76-103): nearest point by clipping,
    farthest corner per axis by the larger |offset| (the lower edge wins ties), nearest = 0 inside the box."""
    o = origin[np.newaxis, :]
    inside = ((o >= lo) & (o <= hi)).all(axis=1)
    near_pt = np.clip(o, lo, hi)
    off_lo, off_hi = o - lo, o - hi
    far_off = np.where(np.abs(off_hi) > np.abs(off_lo), off_hi, off_lo)
    far_pt = o - far_off
    nearest = np.sqrt(((near_pt - o) ** 2).sum(axis=1))
    farthest = np.sqrt(((far_pt - o) ** 2).sum(axis=1))
    nearest[inside] = 0.
    return nearest, farthest


class _SubspacePointSchedulerRectangularND(ActiveScheduler):
    "front expanding from a point of the subspace spanned by the axes NOT listed in `iaxes` (in x units)"

    def __init__(self, all_params, n_steps, point, iaxes):
        super().__init__(all_params, n_steps)
        point, iaxes = np.array(point), list(iaxes)
        if point.ndim != 1:
            raise Exception("ERROR: point.ndim != 1")
        if len(point) > self.xd:
            raise Exception("ERROR: len(point) > self.xd")
        if len(point) + len(iaxes) != self.xd:
            raise Exception("ERROR: len(iaxes) + len(point) != self.xd")
        self.point, self.iaxes = point, iaxes
        d = all_params["static"]["decomposition"]
        self.xmins0, self.xmaxs0 = np.array(d["xmins0"]).copy(), np.array(d["xmaxs0"]).copy()

    def _get_radii(self, point, xmins, xmaxs):
        assert xmins.shape[1] == xmaxs.shape[1] == point.shape[0]
        return _shell_of_boxes(point, xmins, xmaxs)

    def __iter__(self):
        constrained = [ax for ax in range(self.xd) if ax not in self.iaxes]
        nearest, farthest = self._get_radii(self.point, self.xmins0[:, constrained], self.xmaxs0[:, constrained])
        r0, r1 = nearest.min(), farthest.max()
        state = np.zeros(self.m, dtype=int)
        for i in range(self.n_steps):
            r = r0 + (r1 - r0) * (i / (self.n_steps))
            swept = (r >= nearest) & (r < farthest)
            switch_on = swept & (state == 0)
            freeze = ~swept & (state == 1)
            if not (switch_on.any() or freeze.any()):
                yield None
                continue
            state[switch_on] = 1
            state[freeze] = 2
            yield state


class PointSchedulerRectangularND(_SubspacePointSchedulerRectangularND):
    "front expanding from a point of the domain"

    def __init__(self, all_params, n_steps, point):
        xd = all_params["static"]["decomposition"]["xd"]
        if len(point) != xd:
            raise Exception(f"ERROR: point incorrect shape {np.shape(point)}")
        super().__init__(all_params, n_steps, point, iaxes=[])


class LineSchedulerRectangularND(_SubspacePointSchedulerRectangularND):
    "front expanding from a line parallel to axis `iaxis`"

    def __init__(self, all_params, n_steps, point, iaxis):
        xd = all_params["static"]["decomposition"]["xd"]
        if xd < 2:
            raise Exception("ERROR: requires nd >=2")
        if len(point) != xd - 1:
            raise Exception(f"ERROR: point incorrect shape {np.shape(point)}")
        super().__init__(all_params, n_steps, point, iaxes=[iaxis])


class PlaneSchedulerRectangularND(_SubspacePointSchedulerRectangularND):
    "front expanding from a plane spanned by the two axes `iaxes`"

    def __init__(self, all_params, n_steps, point, iaxes):
        xd = all_params["static"]["decomposition"]["xd"]
        if xd < 3:
            raise Exception("ERROR: requires nd >=3")
        if len(point) != xd - 2:
            raise Exception(f"ERROR: point incorrect shape {np.shape(point)}")
        super().__init__(all_params, n_steps, point, iaxes=iaxes)
