"""
Generates the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE'S OWN SOURCE in this container
(/root/reference is read-only and does not travel to the GPU box, so the outputs are committed; this script is the
record of how they were made).  Nothing is copied from the reference: its modules are loaded from where they lie.

  schedulers.npz : fbpinns/schedulers.py imports only numpy -> loaded directly with importlib.
  refmodel_*.npz : see make_golden_shim.py (reference model code executed under a numpy-backed jax shim).

Run:  python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/fbpinns"
sys.path.insert(0, ROOT)


def load_ref_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden_schedulers():
    ref = load_ref_module("ref_schedulers", os.path.join(REF, "schedulers.py"))
    from fbpinns_b200 import configs
    from oracle import ref_takes
    out = {}

    def record(tag, sch):
        steps, states = [], []
        for i, a in enumerate(sch):
            if a is not None:
                steps.append(i)
                states.append(np.array(a).copy())
        out[tag + "_steps"] = np.array(steps)
        out[tag + "_states"] = np.array(states)

    c = configs.cfg3_burgers(n_sub=(6, 5), n_pts=(10, 10))
    d = ref_takes.rectangular_init_params(**c.decomposition_init_kwargs)          # float64 xmins0/xmaxs0
    ap = {"static": {"decomposition": d}}
    record("line", ref.LineSchedulerRectangularND(ap, 50, point=[0.], iaxis=0))
    record("point", ref.PointSchedulerRectangularND(ap, 37, point=np.array([0.3, 0.1])))
    record("all", ref.AllActiveSchedulerND(ap, 5))
    c3 = configs.cfg4_wave3d(n_sub=(3, 4, 5), n_pts=(4, 4, 4))
    d3 = ref_takes.rectangular_init_params(**c3.decomposition_init_kwargs)
    record("plane", ref.PlaneSchedulerRectangularND({"static": {"decomposition": d3}}, 30, point=[0.], iaxes=[0, 1]))
    np.savez(os.path.join(HERE, "schedulers.npz"), **out)
    print("schedulers.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    golden_schedulers()
