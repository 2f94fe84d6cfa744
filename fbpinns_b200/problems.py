"""
PDE problems (API of fbpinns/problems.py:19-62).  `loss_fn` and `constraining_fn` are user plug-in code that
stays in the host framework on the returned ujs — torch here (the reference's are jax.numpy); they are the
boundary of the hot path, not part of it.  The classes below restate the reference problems named by
BASELINE.json plus the synthetic Poisson2D scale-sweep problem.  `exact_solution` is only a test metric: it is
provided where it is closed-form and raises NotImplementedError where the reference calls its finite-difference
/ Burgers quadrature generators (fbpinns/traditional_solutions/, out of scope).
"""
import math

import numpy as np
import torch


class Problem:
    """Base problem class (fbpinns/problems.py:19-62)."""

    @staticmethod
    def init_params(*args):
        raise NotImplementedError

    @staticmethod
    def sample_constraints(all_params, domain, key, sampler, batch_shapes):
        raise NotImplementedError

    @staticmethod
    def constraining_fn(all_params, x_batch, u):
        return u

    @staticmethod
    def loss_fn(all_params, constraints):
        raise NotImplementedError

    @staticmethod
    def exact_solution(all_params, x_batch, batch_shape=None):
        raise NotImplementedError


def _col(v, like=None):
    t = torch.tensor(v, dtype=torch.float32)
    return t.reshape(-1, 1)


class HarmonicOscillator1D(Problem):
    """m u'' + mu u' + k u = 0, u(0)=1, u'(0)=0  (fbpinns/problems.py:68-146)"""

    @staticmethod
    def init_params(d=2, w0=20):
        mu, k = 2 * d, w0 ** 2
        return {"dims": (1, 1), "d": d, "w0": w0, "mu": mu, "k": k}, {}

    @staticmethod
    def sample_constraints(all_params, domain, key, sampler, batch_shapes):
        x_batch_phys = domain.sample_interior(all_params, key, sampler, batch_shapes[0])
        required_ujs_phys = ((0, ()), (0, (0,)), (0, (0, 0)))
        x_batch_boundary = _col([0.])
        u_boundary = _col([1.])
        ut_boundary = _col([0.])
        required_ujs_boundary = ((0, ()), (0, (0,)))
        return [[x_batch_phys, required_ujs_phys],
                [x_batch_boundary, u_boundary, ut_boundary, required_ujs_boundary]]

    @staticmethod
    def loss_fn(all_params, constraints):
        mu, k = all_params["static"]["problem"]["mu"], all_params["static"]["problem"]["k"]
        _, u, ut, utt = constraints[0]
        phys = torch.mean((utt + mu * ut + k * u) ** 2)
        _, uc, utc, u, ut = constraints[1]
        if len(uc):
            boundary = 1e6 * torch.mean((u - uc) ** 2) + 1e2 * torch.mean((ut - utc) ** 2)
        else:
            boundary = 0
        return phys + boundary

    @staticmethod
    def exact_solution(all_params, x_batch, batch_shape=None):
        d, w0 = all_params["static"]["problem"]["d"], all_params["static"]["problem"]["w0"]
        w = math.sqrt(w0 ** 2 - d ** 2)
        phi = math.atan(-d / w)
        A = 1 / (2 * math.cos(phi))
        return torch.exp(-d * x_batch) * 2 * A * torch.cos(phi + w * x_batch)


class HarmonicOscillator1DHardBC(HarmonicOscillator1D):
    """Hard boundary conditions through the constraining operator (fbpinns/problems.py:149-206)"""

    @staticmethod
    def init_params(d=2, w0=20, sd=0.1):
        mu, k = 2 * d, w0 ** 2
        return {"dims": (1, 1), "d": d, "w0": w0, "mu": mu, "k": k, "sd": sd}, {}

    @staticmethod
    def sample_constraints(all_params, domain, key, sampler, batch_shapes):
        x_batch_phys = domain.sample_interior(all_params, key, sampler, batch_shapes[0])
        return [[x_batch_phys, ((0, ()), (0, (0,)), (0, (0, 0)))], ]

    @staticmethod
    def constraining_fn(all_params, x_batch, u):
        sd = all_params["static"]["problem"]["sd"]
        x = x_batch[:, 0:1]
        return 1 + (torch.tanh(x / sd) ** 2) * u

    @staticmethod
    def loss_fn(all_params, constraints):
        mu, k = all_params["static"]["problem"]["mu"], all_params["static"]["problem"]["k"]
        _, u, ut, utt = constraints[0]
        return torch.mean((utt + mu * ut + k * u) ** 2)


class HarmonicOscillator1DInverse(HarmonicOscillator1D):
    """Inverse problem: learn mu from 13 observations (fbpinns/problems.py:209-271)"""

    @staticmethod
    def init_params(d=2, w0=20):
        mu, k = 2 * d, w0 ** 2
        return ({"dims": (1, 1), "d": d, "w0": w0, "mu_true": mu, "k": k},
                {"mu": torch.tensor(0., dtype=torch.float32)})

    @staticmethod
    def sample_constraints(all_params, domain, key, sampler, batch_shapes):
        x_batch_phys = domain.sample_interior(all_params, key, sampler, batch_shapes[0])
        required_ujs_phys = ((0, ()), (0, (0,)), (0, (0, 0)))
        x_batch_data = torch.linspace(0, 1, 13, dtype=torch.float32).reshape(13, 1)
        u_data = HarmonicOscillator1DInverse.exact_solution(all_params, x_batch_data)
        return [[x_batch_phys, required_ujs_phys], [x_batch_data, u_data, ((0, ()),)]]

    @staticmethod
    def loss_fn(all_params, constraints):
        mu, k = all_params["trainable"]["problem"]["mu"], all_params["static"]["problem"]["k"]
        _, u, ut, utt = constraints[0]
        phys = torch.mean((utt + mu * ut + k * u) ** 2)
        _, uc, u = constraints[1]
        data = 1e6 * torch.mean((u - uc) ** 2)
        return phys + data


class BurgersEquation2D(Problem):
    """u_t + u u_x = nu u_xx on [-1,1]x[0,1], u(x,0) = -sin(pi x), u(+-1,t) = 0 (fbpinns/problems.py:276-343)"""

    @staticmethod
    def init_params(nu=0.01 / math.pi, sd=0.1):
        return {"dims": (1, 2), "nu": nu, "sd": sd}, {}

    @staticmethod
    def sample_constraints(all_params, domain, key, sampler, batch_shapes):
        x_batch_phys = domain.sample_interior(all_params, key, sampler, batch_shapes[0])
        return [[x_batch_phys, ((0, ()), (0, (0,)), (0, (1,)), (0, (0, 0)))], ]

    @staticmethod
    def constraining_fn(all_params, x_batch, u):
        sd = all_params["static"]["problem"]["sd"]
        x, t = x_batch[:, 0:1], x_batch[:, 1:2]
        th = torch.tanh
        return th((x + 1) / sd) * th((1 - x) / sd) * th((t - 0) / sd) * u - torch.sin(math.pi * x)

    @staticmethod
    def loss_fn(all_params, constraints):
        nu = all_params["static"]["problem"]["nu"]
        _, u, ux, ut, uxx = constraints[0]
        phys = ut + (u * ux) - (nu * uxx)
        return torch.mean(phys ** 2)


class WaveEquationConstantVelocity3D(Problem):
    """(2+1)D wave equation u_xx + u_yy - u_tt/c^2 = 0 with Gaussian initial wavefield
    (fbpinns/problems.py:346-481)"""

    @staticmethod
    def init_params(c0=1, source=np.array([[0., 0., 0.2, 1.]])):
        return {"dims": (1, 3), "c0": c0, "c_fn": WaveEquationConstantVelocity3D.c_fn,
                "source": torch.tensor(np.asarray(source), dtype=torch.float32)}, {}

    @staticmethod
    def sample_constraints(all_params, domain, key, sampler, batch_shapes):
        x_batch_phys = domain.sample_interior(all_params, key, sampler, batch_shapes[0])
        return [[x_batch_phys, ((0, (0, 0)), (0, (1, 1)), (0, (2, 2)))], ]

    @staticmethod
    def constraining_fn(all_params, x_batch, u):
        params = all_params["static"]["problem"]
        c0, source = params["c0"], params["source"].to(x_batch.device)
        x, t = x_batch[:, 0:2], x_batch[:, 2:3]
        p = source.unsqueeze(1)                                     # (k, 1, 4)
        xx = x.unsqueeze(0)                                         # (1, n, 2)
        f = (p[:, :, 3:4] * torch.exp(-0.5 * ((xx - p[:, :, 0:2]) ** 2).sum(2, keepdim=True) / (p[:, :, 2:3] ** 2))).sum(0)
        t1 = float(source[:, 2].min()) / c0
        f = torch.exp(-0.5 * (1.5 * t / t1) ** 2) * f
        tt = torch.tanh(2.5 * t / t1) ** 2
        return f + tt * u

    @staticmethod
    def loss_fn(all_params, constraints):
        c_fn = all_params["static"]["problem"]["c_fn"]
        x_batch, uxx, uyy, utt = constraints[0]
        phys = (uxx + uyy) - (1 / c_fn(all_params, x_batch) ** 2) * utt
        return torch.mean(phys ** 2)

    @staticmethod
    def c_fn(all_params, x_batch):
        c0 = all_params["static"]["problem"]["c0"]
        return torch.tensor([[c0]], dtype=torch.float32, device=x_batch.device)


class WaveEquationGaussianVelocity3D(WaveEquationConstantVelocity3D):
    """Gaussian-mixture velocity model (fbpinns/problems.py:484-522)"""

    @staticmethod
    def init_params(c0=1, source=np.array([[0., 0., 0.2, 1.]]), mixture=np.array([[0.5, 0.5, 1., 0.2]])):
        return {"dims": (1, 3), "c0": c0, "c_fn": WaveEquationGaussianVelocity3D.c_fn,
                "source": torch.tensor(np.asarray(source), dtype=torch.float32),
                "mixture": torch.tensor(np.asarray(mixture), dtype=torch.float32)}, {}

    @staticmethod
    def c_fn(all_params, x_batch):
        c0, mixture = all_params["static"]["problem"]["c0"], all_params["static"]["problem"]["mixture"].to(x_batch.device)
        x = x_batch[:, 0:2]
        p = mixture.unsqueeze(1)
        xx = x.unsqueeze(0)
        f = (p[:, :, 3:4] * torch.exp(-0.5 * ((xx - p[:, :, 0:2]) ** 2).sum(2, keepdim=True) / (p[:, :, 2:3] ** 2))).sum(0)
        return c0 + f


class Poisson2D(Problem):
    """Synthetic scale-sweep problem of BASELINE.json config 5 (not in the reference; written against its Problem
    API): -(u_xx + u_yy) = f on [0,1]^2, u = 0 on the boundary imposed by a hard-BC constraining operator,
    manufactured solution u* = sin(a pi x) sin(b pi y)."""

    @staticmethod
    def init_params(a=4, b=4, sd=0.1):
        return {"dims": (1, 2), "a": a, "b": b, "sd": sd}, {}

    @staticmethod
    def source(all_params, x_batch):
        "f of -(u_xx + u_yy) = f for the manufactured solution"
        a, b = all_params["static"]["problem"]["a"], all_params["static"]["problem"]["b"]
        x, y = x_batch[:, 0:1], x_batch[:, 1:2]
        return (a * a + b * b) * math.pi ** 2 * torch.sin(a * math.pi * x) * torch.sin(b * math.pi * y)

    @staticmethod
    def sample_constraints(all_params, domain, key, sampler, batch_shapes):
        # the source term is a constraining value carried with the points (like u_boundary / u_data of the reference
        # problems, fbpinns/problems.py:106-113, 250-256): static, so it is not recomputed every step
        x_batch_phys = domain.sample_interior(all_params, key, sampler, batch_shapes[0])
        f_phys = Poisson2D.source(all_params, x_batch_phys)
        return [[x_batch_phys, f_phys, ((0, (0, 0)), (0, (1, 1)))], ]

    @staticmethod
    def constraining_fn(all_params, x_batch, u):
        sd = all_params["static"]["problem"]["sd"]
        x, y = x_batch[:, 0:1], x_batch[:, 1:2]
        th = torch.tanh
        return th(x / sd) * th((1 - x) / sd) * th(y / sd) * th((1 - y) / sd) * u

    @staticmethod
    def loss_fn(all_params, constraints):
        x_batch, f, uxx, uyy = constraints[0]
        return torch.mean((uxx + uyy + f) ** 2)

    @staticmethod
    def exact_solution(all_params, x_batch, batch_shape=None):
        a, b = all_params["static"]["problem"]["a"], all_params["static"]["problem"]["b"]
        return torch.sin(a * math.pi * x_batch[:, 0:1]) * torch.sin(b * math.pi * x_batch[:, 1:2])
