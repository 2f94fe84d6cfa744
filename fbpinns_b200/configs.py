"""
The BASELINE.json configurations as concrete `Constants`, plus reduced variants used by the parity
tests.  `scale` < 1 shrinks the subdomain grid and the collocation grid together so that points per subdomain
stay the same as in the full configuration.
"""
import numpy as np

from . import domains, problems, decompositions, networks, schedulers
from .constants import Constants, get_subdomain_ws


def cfg1_harmonic_oscillator(n_sub=15, n_pts=200, layer_sizes=(1, 32, 1), **kw):
    "HarmonicOscillator1D (d=2, w0=80), 15 subdomains of width 0.15, FCN [1,32,1], 200 points (README.md:123-148)"
    xs = [np.linspace(0, 1, n_sub)]
    return Constants(
        run="cfg1", domain=domains.RectangularDomainND,
        domain_init_kwargs=dict(xmin=np.array([0.]), xmax=np.array([1.])),
        problem=problems.HarmonicOscillator1D, problem_init_kwargs=dict(d=2, w0=80),
        decomposition=decompositions.RectangularDecompositionND,
        decomposition_init_kwargs=dict(subdomain_xs=xs, subdomain_ws=[0.15 * np.ones((n_sub,))], unnorm=(0., 1.)),
        network=networks.FCN, network_init_kwargs=dict(layer_sizes=list(layer_sizes)),
        ns=((n_pts,),), n_test=(500,), **kw)


def cfg2_harmonic_oscillator_inverse(n_sub=30, n_pts=200, layer_sizes=(1, 32, 32, 1), **kw):
    "HarmonicOscillator1DInverse with trainable mu, 30 subdomains (width 2.1 x spacing), FCN [1,32,32,1]"
    xs = [np.linspace(0, 1, n_sub)]
    return Constants(
        run="cfg2", domain=domains.RectangularDomainND,
        domain_init_kwargs=dict(xmin=np.array([0.]), xmax=np.array([1.])),
        problem=problems.HarmonicOscillator1DInverse, problem_init_kwargs=dict(d=2, w0=20),
        decomposition=decompositions.RectangularDecompositionND,
        decomposition_init_kwargs=dict(subdomain_xs=xs, subdomain_ws=get_subdomain_ws(xs, 2.1), unnorm=(0., 1.)),
        network=networks.FCN, network_init_kwargs=dict(layer_sizes=list(layer_sizes)),
        ns=((n_pts,),), n_test=(500,), **kw)


def cfg3_burgers(n_sub=(15, 15), n_pts=(200, 200), layer_sizes=(2, 16, 1), line_scheduler=True, n_steps=15000, **kw):
    "BurgersEquation2D, 15x15 subdomains (w = 2.9 dx), unnorm (0,3), 200x200 points, LineScheduler along t"
    xs = [np.linspace(-1, 1, n_sub[0]), np.linspace(0, 1, n_sub[1])]
    sched = dict(scheduler=schedulers.LineSchedulerRectangularND,
                 scheduler_kwargs=dict(point=[0.], iaxis=0)) if line_scheduler else {}
    return Constants(
        run="cfg3", domain=domains.RectangularDomainND,
        domain_init_kwargs=dict(xmin=np.array([-1., 0.]), xmax=np.array([1., 1.])),
        problem=problems.BurgersEquation2D, problem_init_kwargs=dict(),
        decomposition=decompositions.RectangularDecompositionND,
        decomposition_init_kwargs=dict(subdomain_xs=xs, subdomain_ws=get_subdomain_ws(xs, 2.9), unnorm=(0., 3.)),
        network=networks.FCN, network_init_kwargs=dict(layer_sizes=list(layer_sizes)),
        ns=(tuple(n_pts),), n_test=(100, 100), n_steps=n_steps, **sched, **kw)


def cfg4_wave3d(n_sub=(10, 10, 10), n_pts=(64, 64, 64), layer_sizes=(3, 64, 64, 1), **kw):
    "WaveEquationGaussianVelocity3D, 10^3 subdomains (w = 2.9 dx), FCN [3,64,64,1], second-order ujs"
    xs = [np.linspace(-1, 1, n_sub[0]), np.linspace(-1, 1, n_sub[1]), np.linspace(0, 1, n_sub[2])]
    return Constants(
        run="cfg4", domain=domains.RectangularDomainND,
        domain_init_kwargs=dict(xmin=np.array([-1., -1., 0.]), xmax=np.array([1., 1., 1.])),
        problem=problems.WaveEquationGaussianVelocity3D, problem_init_kwargs=dict(),
        decomposition=decompositions.RectangularDecompositionND,
        decomposition_init_kwargs=dict(subdomain_xs=xs, subdomain_ws=get_subdomain_ws(xs, 2.9), unnorm=(0., 1.)),
        network=networks.FCN, network_init_kwargs=dict(layer_sizes=list(layer_sizes)),
        ns=(tuple(n_pts),), n_test=(20, 20, 10), **kw)


def cfg5_poisson(n_sub=(64, 64), n_pts=(1024, 1024), layer_sizes=(2, 32, 32, 1), **kw):
    "Synthetic 2D Poisson scale sweep: 64x64 = 4096 subdomains (w = 2.9 dx), 1024x1024 collocation grid"
    xs = [np.linspace(0, 1, n_sub[0]), np.linspace(0, 1, n_sub[1])]
    return Constants(
        run="cfg5", domain=domains.RectangularDomainND,
        domain_init_kwargs=dict(xmin=np.array([0., 0.]), xmax=np.array([1., 1.])),
        problem=problems.Poisson2D, problem_init_kwargs=dict(),
        decomposition=decompositions.RectangularDecompositionND,
        decomposition_init_kwargs=dict(subdomain_xs=xs, subdomain_ws=get_subdomain_ws(xs, 2.9), unnorm=(0., 1.)),
        network=networks.FCN, network_init_kwargs=dict(layer_sizes=list(layer_sizes)),
        ns=(tuple(n_pts),), n_test=(128, 128), **kw)


CONFIGS = {"cfg1": cfg1_harmonic_oscillator, "cfg2": cfg2_harmonic_oscillator_inverse, "cfg3": cfg3_burgers,
           "cfg4": cfg4_wave3d, "cfg5": cfg5_poisson}

# reduced variants (same points per subdomain as the full configuration) for parity tests against the oracle
SMALL = {
    "cfg1": dict(),
    "cfg2": dict(),
    "cfg3": dict(n_sub=(5, 5), n_pts=(66, 66)),
    "cfg4": dict(n_sub=(3, 3, 3), n_pts=(18, 18, 18), layer_sizes=(3, 16, 16, 1)),
    "cfg5": dict(n_sub=(6, 6), n_pts=(96, 96)),
}
