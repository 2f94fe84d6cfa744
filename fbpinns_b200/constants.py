"""
Constants object: problem setup + hyperparameters (same attributes, defaults and keyed-override behaviour as
fbpinns/constants.py:27-97 and fbpinns/constants_base.py:12-55; unknown keys are rejected).
"""
import os
import pickle
import socket

import numpy as np

from . import domains, problems, decompositions, networks, schedulers


def get_subdomain_ws(subdomain_xs, width):
    """fbpinns/constants.py:21-22"""
    return [width * np.min(np.diff(x)) * np.ones_like(x) for x in subdomain_xs]


class ConstantsBase:
    def __getitem__(self, key):
        if key not in vars(self):
            raise KeyError(f'key "{key}" not defined in class')
        return getattr(self, key)

    def __setitem__(self, key, item):
        if key not in vars(self):
            raise KeyError(f'key "{key}" not defined in class')
        setattr(self, key, item)

    def __str__(self):
        s = repr(self) + "\n"
        for k in vars(self):
            s += f"{k}: {self[k]}\n"
        return s

    @property
    def summary_out_dir(self):
        return f"results/summaries/{self.run}/"

    @property
    def model_out_dir(self):
        return f"results/models/{self.run}/"

    def get_outdirs(self):
        for d in (self.summary_out_dir, self.model_out_dir):
            os.makedirs(d, exist_ok=True)

    def save_constants_file(self):
        with open(self.summary_out_dir + f"constants_{self.run}.txt", "w") as f:
            for k in vars(self):
                f.write(f"{k}: {self[k]}\n")
        with open(self.summary_out_dir + f"constants_{self.run}.pickle", "wb") as f:
            pickle.dump(vars(self), f)

    @property
    def constants_file(self):
        return self.summary_out_dir + f"constants_{self.run}.pickle"


class Constants(ConstantsBase):

    def __init__(self, **kwargs):
        self.run = "test"

        self.domain = domains.RectangularDomainND
        self.domain_init_kwargs = dict(xmin=np.array([0.]), xmax=np.array([1.]))

        self.problem = problems.HarmonicOscillator1D
        self.problem_init_kwargs = dict(d=2, w0=20)

        subdomain_xs = [np.linspace(0, 1, 5)]
        subdomain_ws = get_subdomain_ws(subdomain_xs, 2.99)
        self.decomposition = decompositions.RectangularDecompositionND
        self.decomposition_init_kwargs = dict(subdomain_xs=subdomain_xs, subdomain_ws=subdomain_ws, unnorm=(0., 1.))

        self.network = networks.FCN
        self.network_init_kwargs = dict(layer_sizes=[1, 32, 1])

        self.n_steps = 15000
        self.scheduler = schedulers.AllActiveSchedulerND
        self.scheduler_kwargs = dict()

        self.ns = ((60,),)
        self.n_test = (200,)
        self.sampler = "grid"
        self.optimiser = "adam"          # the reference hard-codes optax.adam (fbpinns/trainers.py:430)
        self.optimiser_kwargs = dict(learning_rate=1e-3)
        self.seed = 0

        self.summary_freq = 1000
        self.test_freq = 1000
        self.model_save_freq = 10000
        self.show_figures = False
        self.save_figures = False
        self.clear_output = False

        # B200 engine options (not in the reference)
        self.device = "cuda:0"
        self.use_cuda_graph = True       # capture the whole step between active-set changes
        self.kernel = "auto"             # "auto" | "generic" | "tiled" | "tensor" | "tensor-full" (tcgen05 where an instance exists)
        self.save_models = False         # write model_{i:08d}.jax checkpoints (reference format) every model_save_freq steps
        self.init_prng = "numpy"         # "numpy" (default_rng(seed)) | "jax" (the reference's threefry key sequence, util/jax_prng.py)

        self.hostname = socket.gethostname().lower()

        for key in kwargs.keys():
            self[key] = kwargs[key]
