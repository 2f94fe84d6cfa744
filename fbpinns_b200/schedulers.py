"""
Active schedulers (API and semantics of fbpinns/schedulers.py:16-161): iterables yielding, per training step,
either None (active set unchanged) or an int array of length m with 0 = inactive, 1 = active, 2 = fixed.
Pure float64 numpy on the host, exactly as in the reference — the output is consumed bit-exactly by the
trainer's active-set algebra; it is ~100 lines of host logic, not a kernel target.
"""
import numpy as np


class ActiveScheduler:
    """Base scheduler class"""

    def __init__(self, all_params, n_steps):
        self.n_steps = n_steps
        self.m = all_params["static"]["decomposition"]["m"]
        self.xd = all_params["static"]["decomposition"]["xd"]

    def __len__(self):
        return self.n_steps

    def __iter__(self):
        raise NotImplementedError


class AllActiveSchedulerND(ActiveScheduler):
    "All models are active and training all of the time"

    def __iter__(self):
        for i in range(self.n_steps):
            yield np.ones(self.m, dtype=int) if i == 0 else None


class _SubspacePointSchedulerRectangularND(ActiveScheduler):
    "Slowly expands radially outwards from a point in a subspace of a rectangular domain (in x units)"

    def __init__(self, all_params, n_steps, point, iaxes):
        super().__init__(all_params, n_steps)
        point = np.array(point)
        iaxes = list(iaxes)
        if point.ndim != 1:
            raise Exception("ERROR: point.ndim != 1")
        if len(point) > self.xd:
            raise Exception("ERROR: len(point) > self.xd")
        if len(iaxes) + len(point) != self.xd:
            raise Exception("ERROR: len(iaxes) + len(point) != self.xd")
        self.point = point
        self.iaxes = iaxes
        self.xmins0 = np.array(all_params["static"]["decomposition"]["xmins0"]).copy()
        self.xmaxs0 = np.array(all_params["static"]["decomposition"]["xmaxs0"]).copy()

    def _get_radii(self, point, xmins, xmaxs):
        "Nearest / farthest distance from the point to each hyperrectangle (0 nearest if inside)"
        assert xmins.shape[1] == xmaxs.shape[1] == point.shape[0]
        pt = point[None, :]
        inside = np.all((pt >= xmins) & (pt <= xmaxs), axis=1)
        nearest = np.clip(pt, xmins, xmaxs)
        d_lo, d_hi = pt - xmins, pt - xmaxs
        use_hi = np.abs(d_hi) > np.abs(d_lo)            # argmax over [d_lo, d_hi] picks d_lo on ties
        farthest = pt - np.where(use_hi, d_hi, d_lo)
        rmin = np.sqrt(np.sum((nearest - pt) ** 2, axis=1))
        rmax = np.sqrt(np.sum((farthest - pt) ** 2, axis=1))
        rmin[inside] = 0.
        return rmin, rmax

    def __iter__(self):
        ic = [i for i in range(self.xd) if i not in self.iaxes]
        rmin, rmax = self._get_radii(self.point, self.xmins0[:, ic], self.xmaxs0[:, ic])
        r_lo, r_hi = rmin.min(), rmax.max()
        active = np.zeros(self.m, dtype=int)
        for i in range(self.n_steps):
            rt = r_lo + (r_hi - r_lo) * (i / (self.n_steps))
            in_ring = (rt >= rmin) & (rt < rmax)
            to_active = (active == 0) & in_ring
            to_fixed = (active == 1) & (~in_ring)
            if to_active.any() or to_fixed.any():
                active[to_active] = 1
                active[to_fixed] = 2
                yield active
            else:
                yield None


class PointSchedulerRectangularND(_SubspacePointSchedulerRectangularND):
    "Slowly expands outwards from a point in the domain (in x units)"

    def __init__(self, all_params, n_steps, point):
        xd = all_params["static"]["decomposition"]["xd"]
        if len(point) != xd:
            raise Exception(f"ERROR: point incorrect shape {np.shape(point)}")
        super().__init__(all_params, n_steps, point, iaxes=[])


class LineSchedulerRectangularND(_SubspacePointSchedulerRectangularND):
    "Slowly expands outwards from a line in the domain (in x units)"

    def __init__(self, all_params, n_steps, point, iaxis):
        xd = all_params["static"]["decomposition"]["xd"]
        if xd < 2:
            raise Exception("ERROR: requires nd >=2")
        if len(point) != xd - 1:
            raise Exception(f"ERROR: point incorrect shape {np.shape(point)}")
        super().__init__(all_params, n_steps, point, iaxes=[iaxis])


class PlaneSchedulerRectangularND(_SubspacePointSchedulerRectangularND):
    "Slowly expands outwards from a plane in the domain (in x units)"

    def __init__(self, all_params, n_steps, point, iaxes):
        xd = all_params["static"]["decomposition"]["xd"]
        if xd < 3:
            raise Exception("ERROR: requires nd >=3")
        if len(point) != xd - 2:
            raise Exception(f"ERROR: point incorrect shape {np.shape(point)}")
        super().__init__(all_params, n_steps, point, iaxes=iaxes)
