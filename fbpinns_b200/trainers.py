"""
FBPINN trainer on the B200 engine: same public surface as the reference's trainer
(`FBPINNTrainer(c).train() -> all_params`, fbpinns/trainers.py:479-774) and the same loop structure
(:631-677): on every active-set change rebuild the update inputs, every step call one update.

What differs from the reference by design:
  * the update is not an XLA executable: it is a fixed sequence of hand-written CUDA kernels
    (fbp_forward / fbp_reduce_* / fbp_backward / fbp_adam_step) linked to the user's torch `loss_fn` /
    `constraining_fn` by an autograd tape, captured ONCE per active set into a CUDA graph — an active-set
    change costs an index rebuild on the device and a graph re-capture, never a compilation;
  * parameters live in one packed (m, P) buffer; "cutting" active / fixed parameter trees
    (fbpinns/trainers.py:394-417, 528-531) is an index array (all_ims), not a copy, and Adam state is never
    cut or merged: rows of inactive subdomains are simply not touched.
Out of scope: PINNTrainer, plotting, tensorboard, model checkpoint files.
"""

import os
import time

import numpy as np
import torch

from . import networks, decompositions
from .engine import (Plan, DeviceTakes, ConstraintEvaluator, PackedAdam, pack_params, unpack_params,
                     subdomain_sum, gather_rows, nonzero_i32, device_info)
from .jets import JetSpec, AffineConstraining, get_jmaps  # noqa: F401  (get_jmaps: reference API name)
from .util.logger import logger


# --------------------------------------------------------------------------------------------------- active-set algebra

def active_set_algebra(active, training_model_count):
    """Mask algebra of get_inputs (fbpinns/trainers.py:347-367) on host int arrays of length m:
    inactive (0) -> active (1); models holding no training point -> discarded; returns
    (active, active_ims, fixed_ims, all_ims, pos_of_model)."""
    active = np.array(active).copy()
    m = active.shape[0]
    assert np.isin(active, [0, 1, 2]).all()
    active[active == 0] = 1
    mask = (np.asarray(training_model_count) > 0).astype(active.dtype)
    active = active * mask
    ims = np.arange(m)
    active_ims = ims[active == 1]
    fixed_ims = ims[active == 2]
    all_ims = np.concatenate([active_ims, fixed_ims])
    pos = -np.ones(m, dtype=np.int32)
    pos[all_ims] = np.arange(len(all_ims), dtype=np.int32)
    return active, active_ims.astype(np.int32), fixed_ims.astype(np.int32), all_ims.astype(np.int32), pos


class UpdateInputs:
    """Result of an active-set change: the device-side equivalent of what _get_update_inputs returns
    (fbpinns/trainers.py:509-575)."""
    pass


def get_update_inputs(active, all_params, dd, x_batch_global, constraints_global, constraint_offsets,
                      jets, layer_sizes, kernel="auto", activation="tanh", plans=None):
    """Device implementation of FBPINNTrainer._get_x_batch + _get_update_inputs (index side).
    constraints_global[ic] = list of per-point CUDA tensors of constraint ic (first is its x_batch)."""
    m = dd.m
    active = np.array(active).copy()
    assert np.isin(active, [0, 1, 2]).all()
    assert active.shape == (m,)
    dev = dd.device

    # --- _get_x_batch (fbpinns/trainers.py:482-507): points inside >= 1 scheduler-active model
    ims1 = torch.as_tensor(np.arange(m, dtype=np.int32)[active == 1], dtype=torch.int32, device=dev)
    pt_count, mc1 = dd.inside_count(x_batch_global, models=ims1)
    training_ips = nonzero_i32(pt_count)
    d_stat = float(mc1.double().mean().item() ** (1 / dd.xd)) if ims1.numel() else float("nan")
    x_batch = gather_rows(x_batch_global, training_ips)

    # constraint membership: training_ips is ascending, constraints are contiguous ranges of the global batch
    ips_host = training_ips.cpu().numpy().astype(np.int64)
    sizes = [c_[0].shape[0] for c_ in constraints_global]
    bounds = np.searchsorted(ips_host, np.concatenate([constraint_offsets, [constraint_offsets[-1] + sizes[-1]]]))
    constraint_ips, constraints, local_idx = [], [], []
    for ic in range(len(constraints_global)):
        a, b = int(bounds[ic]), int(bounds[ic + 1])
        constraint_ips.append(np.arange(a, b))
        local = (training_ips[a:b] - int(constraint_offsets[ic])).contiguous()
        local_idx.append(local)
        constraints.append([gather_rows(c_, local) if c_.dtype == torch.float32 else c_[local.long()]
                            for c_ in constraints_global[ic]])

    # --- get_inputs (fbpinns/trainers.py:332-391): which models hold training points, active-set algebra
    _, model_count = dd.inside_count(x_batch, models=None)
    active2, active_ims, fixed_ims, all_ims, pos = active_set_algebra(active, model_count.cpu().numpy())

    out = UpdateInputs()
    out.active, out.active_ims, out.fixed_ims, out.all_ims, out.pos_of_model = active2, active_ims, fixed_ims, all_ims, pos
    out.x_batch, out.constraints, out.training_ips, out.constraint_ips, out.d = x_batch, constraints, training_ips, constraint_ips, d_stat
    out.takess, out.evaluators = [], []
    out.constraint_local_idx = local_idx            # rows of each constraint's global arrays that are training points now
    for ic, con in enumerate(constraints):
        # plans are static (network shape + jet set): the trainer hands in its own so that an active-set change creates none
        plan = plans[ic] if plans is not None else Plan(layer_sizes, jets[ic], kernel=kernel, activation=activation)
        takes = DeviceTakes(dd, con[0], pos, all_ims, len(active_ims), tile_points=plan.tile_points)
        out.takess.append(takes)
        out.evaluators.append(ConstraintEvaluator(plan, takes, con[0], dd))
    return out


# --------------------------------------------------------------------------------------------------- the update step

class UpdateStep:
    """One FBPINN_update (fbpinns/trainers.py:285-296) for a fixed active set, optionally captured as a CUDA graph."""

    def __init__(self, inputs, params, adam, all_params, prob_flat, problem, use_cuda_graph=True, affine_global=None):
        self.inp, self.params, self.adam = inputs, params, adam
        self.all_params, self.prob_flat, self.problem = all_params, prob_flat, problem
        dev = params.device
        self.grads = torch.zeros((max(len(inputs.active_ims), 1), params.shape[1]), dtype=torch.float32, device=dev)
        self.active_ims_dev = torch.as_tensor(inputs.active_ims, dtype=torch.int32, device=dev)
        self.hook = torch.zeros((), dtype=torch.float32, device=dev, requires_grad=True)
        self.loss_out = torch.zeros((), dtype=torch.float32, device=dev)
        from .problems import Problem
        self.has_constraining = problem.constraining_fn is not Problem.constraining_fn
        self.prob_keys = list(all_params["trainable"].get("problem", {}).keys()) if prob_flat is not None else []
        self.prob_shapes = [tuple(all_params["trainable"]["problem"][k].shape) for k in self.prob_keys]
        self.use_graph = use_cuda_graph
        self.graph = None
        self.n_eager = 0
        self.kernel_launches_per_step = None
        # constraining operators that are affine in u (all hard-BC ansatzes of the reference problems) get static
        # coefficient jets, computed once here; anything else goes through the generic nested-jvp path every step
        # `affine_global[ic]` (FBPINNTrainer.setup): the same jets computed ONCE for all points of the constraint; an
        # active-set change then only gathers the rows of the current training points
        self.affine = []
        for ic, (ev, con) in enumerate(zip(inputs.evaluators, inputs.constraints)):
            base = ev.ev if hasattr(ev, "ev") else ev             # sharded evaluators wrap the local one
            aff = None
            if self.has_constraining and prob_flat is None and base.takes.npou == 1:
                rows = getattr(inputs, "kernel_point_rows", None)
                if affine_global is not None and rows is not None:
                    ag = affine_global[ic]
                    aff = None if ag is None else AffineConstraining(ag.jet, gather_rows(ag.Aj, rows[ic]), gather_rows(ag.Bj, rows[ic]))
                else:
                    aff = AffineConstraining.build(base.plan.jet, base.x, problem.constraining_fn, all_params)
            base.set_affine(aff)            # fused into the reduce kernels (forward Leibniz rule and its transpose)
            self.affine.append(aff)
        # a single (unsharded) constraint whose work list has one item per active subdomain: the reverse kernel writes the rows
        # of `grads` itself — no zero fill, no partial buffer, no reduction pass (FBP_BWD_DIRECT; FBP_DIRECT_GRADS=0 turns it off)
        evs = inputs.evaluators
        self.direct_grads = bool(len(evs) == 1 and not hasattr(evs[0], "ev") and hasattr(evs[0], "supports_direct_grads")
                                 and evs[0].supports_direct_grads() and os.environ.get("FBP_DIRECT_GRADS", "1") != "0")
        for e_ in evs:
            if hasattr(e_, "supports_direct_grads"):
                e_.direct_grads = self.direct_grads

    def _refresh_problem_views(self):
        """Problem trainables live in one flat storage buffer (updated in place by Adam).  Every step aliases it with a
        FRESH leaf created on the current stream — a persistent leaf would keep an AccumulateGrad node bound to the
        stream it was first used on, which breaks CUDA-graph capture — and rebuilds the views user code reads."""
        self._pf = None
        if self.prob_flat is None:
            return
        self._pf = self.prob_flat.detach().requires_grad_(True)
        off = 0
        for k, shp in zip(self.prob_keys, self.prob_shapes):
            nel = int(np.prod(shp)) if len(shp) else 1
            self.all_params["trainable"]["problem"][k] = self._pf[off:off + nel].view(shp)
            off += nel

    def problem_grad(self):
        "gradient of the flat problem-parameter buffer from the last backward (zeros if unused)"
        if self._pf is None or self.prob_flat.numel() == 0:
            return None
        return self._pf.grad if self._pf.grad is not None else torch.zeros_like(self.prob_flat)

    def forward_loss(self):
        self._refresh_problem_views()
        cons = []
        for ev, con, aff in zip(self.inp.evaluators, self.inp.constraints, self.affine):
            ujets = subdomain_sum(ev, self.params, self.grads, self.hook)
            jet = ev.plan.jet
            if aff is not None:
                ujs = jet.ujs_plain(ujets)            # the kernels already returned the constrained jets
            elif self.has_constraining:
                ujs = jet.ujs_constrained(ujets, con[0], self.problem.constraining_fn, self.all_params)
            else:
                ujs = jet.ujs_plain(ujets)
            cons.append(list(con) + ujs)
        return self.problem.loss_fn(self.all_params, cons)

    def _eager(self):
        if not self.direct_grads:
            self.grads.zero_()
        self.hook.grad = None
        loss = self.forward_loss()
        loss.backward()
        pg = self.problem_grad()
        with torch.no_grad():
            self.adam.step(self.params, self.grads, self.active_ims_dev,
                           self.prob_flat if pg is not None else None, pg)
            self.loss_out.copy_(loss.detach())
        return self.loss_out

    def __call__(self):
        """Runs one update; returns the (device, scalar) loss evaluated BEFORE the update, like value_and_grad."""
        if not self.use_graph:
            return self._eager()
        if self.graph is None:
            if self.n_eager < 3:                      # warm-up steps are real training steps
                self.n_eager += 1
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._eager()
                torch.cuda.current_stream().wait_stream(s)
                return self.loss_out
            self.hook.grad = None
            from . import _lib
            n0 = _lib.load().fbp_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._eager()
            self.graph = g
            self.kernel_launches_per_step = int(_lib.load().fbp_launch_count() - n0)
        self.graph.replay()
        return self.loss_out


# --------------------------------------------------------------------------------------------------- trainer

class _Trainer:
    "Generic model trainer base class (logging only; the reference's tensorboard / figure / pickle output is out of scope)"

    def __init__(self, c):
        self.c = c
        logger.info(str(c))

    def _print_summary(self, i, loss, rate, start):
        logger.info("[i: %i/%i] loss: %.4f rate: %.1f elapsed: %.2f hr %s" % (
            i, self.c.n_steps, loss, rate, (time.time() - start) / (60 * 60), self.c.run))


class FBPINNTrainer(_Trainer):
    "FBPINN model trainer class"

    def _init_all_params(self):
        c = self.c
        dev = torch.device(c.device)
        all_params = {"static": {}, "trainable": {}}
        domain, problem, decomposition = c.domain, c.problem, c.decomposition
        for tag, cl, kwargs in zip(["domain", "problem", "decomposition"], [domain, problem, decomposition],
                                   [c.domain_init_kwargs, c.problem_init_kwargs, c.decomposition_init_kwargs]):
            ps_ = cl.init_params(**kwargs)
            if ps_[0]:
                all_params["static"][tag] = ps_[0]
            if ps_[1]:
                all_params["trainable"][tag] = ps_[1]
        assert (all_params["static"]["domain"]["xd"] ==
                all_params["static"]["problem"]["dims"][1] ==
                all_params["static"]["decomposition"]["xd"])
        m = all_params["static"]["decomposition"]["m"]
        logger.info(f"Total number of subdomains: {m}")

        network = c.network
        if getattr(network, "ACTIVATION", None) is None or not hasattr(network, "init_params_batched"):
            raise NotImplementedError(f"{network} is not implemented by the B200 kernels (FCN, AdaptiveFCN, SIREN, "
                                      f"AdaptiveSIREN and FourierFCN are)")
        if not issubclass(decomposition, decompositions.RectangularDecompositionND):
            raise NotImplementedError(f"{decomposition} is not implemented by the B200 kernels "
                                      f"(RectangularDecompositionND family with the cosine window only)")
        if getattr(c, "init_prng", "numpy") == "jax":
            # the reference's key derivation (fbpinns/trainers.py:583, 603-605) on the restated threefry generator
            from .util import jax_prng
            rng = jax_prng.PRNGKey(c.seed)
        else:
            rng = np.random.default_rng(c.seed)
        ps_ = network.init_params_batched(rng, m, **c.network_init_kwargs)
        if ps_[0]:
            all_params["static"]["network"] = {"subdomain": ps_[0]}
        if ps_[1]:
            all_params["trainable"]["network"] = {"subdomain": ps_[1]}
        return all_params, dev

    def setup(self):
        """Everything train() does before its loop (fbpinns/trainers.py:580-624): parameters, constraints, jet specs,
        packed parameter / Adam buffers, test points.  Public so that single steps can be driven from outside
        (bench.py, notebooks): setup() -> set_active(active) -> step() ..."""
        c = self.c
        np.random.seed(c.seed)
        all_params, dev = self._init_all_params()
        if dev.type == "cuda":
            # the C ABI launches on torch's CURRENT device and stream: make it the trainer's device
            torch.cuda.set_device(dev)
        domain, problem, decomposition = c.domain, c.problem, c.decomposition
        m = all_params["static"]["decomposition"]["m"]
        ud, xd = all_params["static"]["problem"]["dims"]
        # kernel view of the network: activation tag, layer sizes and leaves (FourierFCN gains its static feature layer)
        self.activation, self.layer_sizes, layers = networks.kernel_layers(c.network, all_params, dev)
        self.dd = decomposition._device(all_params, dev)

        # constraints (fbpinns/trainers.py:436-461)
        key = np.random.default_rng(c.seed + 1)
        constraints_global = problem.sample_constraints(all_params=all_params, domain=domain, key=key,
                                                        sampler=c.sampler, batch_shapes=c.ns)
        for con in constraints_global:
            for c_ in con[:-1]:
                assert c_.shape[0] == con[0].shape[0]
        required_ujss = [con[-1] for con in constraints_global]
        self.constraints_global = [[t.to(dev, torch.float32).contiguous() for t in con[:-1]] for con in constraints_global]
        self.x_batch_global = torch.cat([con[0] for con in self.constraints_global]).contiguous()
        sizes = [con[0].shape[0] for con in self.constraints_global]
        self.constraint_offsets = np.cumsum([0] + sizes[:-1]).astype(np.int64)
        self.jets = [JetSpec(r, xd, ud) for r in required_ujss]
        shard = getattr(self, "shard", None)
        if shard is not None and shard.world > 1 and os.environ.get("FBP_SHARD_BALANCE", "1") == "1":
            # contiguous blocks of subdomains with equal PAIR counts over the full point set (fixed for the whole run, so
            # that parameter and optimiser-state ownership never moves)
            _, mc = self.dd.inside_count(self.x_batch_global)
            shard.balance(mc.cpu().numpy())
        logger.info(f"Total number of constraints: {len(self.constraints_global)}")

        # packed parameters + problem trainables + Adam
        self.value_plan = Plan(self.layer_sizes, JetSpec(tuple((iu, ()) for iu in range(ud)), xd, ud), kernel=c.kernel,
                               activation=self.activation)
        self.params = pack_params(self.value_plan, layers)
        prob_tr = all_params["trainable"].get("problem", {})
        prob_keys = list(prob_tr.keys())
        if prob_keys:
            flat = torch.cat([prob_tr[k].reshape(-1).float() for k in prob_keys]).to(dev)
            self.prob_flat = flat.clone()          # plain storage; every step aliases it with a fresh grad leaf
            off = 0
            for k in prob_keys:
                nel = prob_tr[k].numel()
                all_params["trainable"]["problem"][k] = self.prob_flat[off:off + nel].view(prob_tr[k].shape)
                off += nel
        else:
            self.prob_flat = None
        for tag in ("domain", "problem"):
            st = all_params["static"].get(tag, {})
            for k, v in st.items():
                if torch.is_tensor(v):
                    st[k] = v.to(dev)
        self.adam = PackedAdam(m, self.value_plan.P, 0 if self.prob_flat is None else self.prob_flat.numel(), dev,
                               **c.optimiser_kwargs)
        logger.info(f"Total number of trainable parameters: network: {self.params.numel():,}")

        # test data (fbpinns/trainers.py:463-470, 620-624)
        self.x_batch_test = domain.sample_interior(all_params=all_params, key=None, sampler="grid",
                                                   batch_shape=c.n_test).to(dev)
        try:
            self.u_exact = problem.exact_solution(all_params=all_params, x_batch=self.x_batch_test, batch_shape=c.n_test)
        except NotImplementedError:
            self.u_exact = None
        self._test_eval = None
        self.all_params = all_params
        # static across active-set changes: one plan per constraint, and the jets of an affine constraining operator
        # A(x) u + B(x) at EVERY point of the constraint (an active-set change gathers rows instead of re-deriving them)
        self.plans = [Plan(self.layer_sizes, j, kernel=c.kernel, activation=self.activation) for j in self.jets]
        from .problems import Problem
        self.affine_global = None
        if problem.constraining_fn is not Problem.constraining_fn and self.prob_flat is None and self.dd.npou == 1:
            self.affine_global = [AffineConstraining.build(j, con[0], problem.constraining_fn, all_params)
                                  for j, con in zip(self.jets, self.constraints_global)]
            for ag in self.affine_global:
                if ag is not None:
                    ag.Aj, ag.Bj = ag.Aj.contiguous().float(), ag.Bj.contiguous().float()
        self.n_rebuilds = 0
        self.inputs = self.update = None
        return self

    def set_active(self, active, i=0):
        "Active-set change (fbpinns/trainers.py:634-653): rebuild the update inputs on the device; no compilation."
        c = self.c
        t0 = time.time()
        logger.info(f"[i: {i}/{c.n_steps}] Updating active inputs..")
        shard = getattr(self, "shard", None)
        # release the previous active set's buffers (and its captured graph) first: the caching allocator then hands the same
        # blocks to the new takes / evaluators instead of going to cudaMalloc for a second copy
        self.update = None
        self.inputs = None
        if shard is None or shard.world == 1:
            self.inputs = get_update_inputs(active, self.all_params, self.dd, self.x_batch_global, self.constraints_global,
                                            self.constraint_offsets, self.jets, self.layer_sizes, kernel=c.kernel,
                                            activation=self.activation, plans=self.plans)
            self.inputs.kernel_point_rows = self.inputs.constraint_local_idx
            self.update = UpdateStep(self.inputs, self.params, self.adam, self.all_params, self.prob_flat, c.problem,
                                     c.use_cuda_graph, affine_global=self.affine_global)
        else:
            from . import parallel
            # several constraints or trainables of the problem itself: evaluate loss_fn on the full ujs on every rank
            replicated = len(self.constraints_global) > 1 or self.prob_flat is not None
            self.inputs = parallel.get_update_inputs_sharded(shard, active, self.all_params, self.dd, self.x_batch_global,
                                                             self.constraints_global, self.constraint_offsets, self.jets,
                                                             self.layer_sizes, kernel=c.kernel, activation=self.activation,
                                                             replicated=replicated, plans=self.plans)
            self.update = parallel.make_sharded_update(UpdateStep)(shard, self.inputs, self.params, self.adam,
                                                                   self.all_params, self.prob_flat, c.problem,
                                                                   c.use_cuda_graph, affine_global=self.affine_global)
        self.n_rebuilds += 1
        torch.cuda.synchronize()
        logger.info(f"[i: {i}/{c.n_steps}] Updating active inputs done ({time.time() - t0:.2f} s); "
                    f"average points/dimension in active subdomains: {self.inputs.d:.2f}")
        return self.inputs

    def step(self):
        "One FBPINN_update on the current active set; returns the device scalar loss (no host sync)."
        return self.update()

    def point_buffers(self):
        "device tensors holding the collocation points a step reads (one or two per constraint)"
        bufs = []
        for ev, con in zip(self.inputs.evaluators, self.inputs.constraints):
            x_kernel = ev.ev.x if hasattr(ev, "ev") else ev.x
            bufs.append(x_kernel)
            if con[0].data_ptr() != x_kernel.data_ptr():
                bufs.append(con[0])
        return bufs

    def step_from_host(self, x_host_pinned_list):
        """End-to-end step: copies the collocation points from (pinned) host memory into the device buffers of
        point_buffers(), runs one update and returns the loss as a python float (device->host read)."""
        for buf, xh in zip(self.point_buffers(), x_host_pinned_list):
            buf.copy_(xh, non_blocking=True)
        return float(self.update().item())

    def train(self):
        "Train model"
        c = self.c
        self.setup()
        scheduler = c.scheduler(all_params=self.all_params, n_steps=c.n_steps, **c.scheduler_kwargs)

        # train loop (fbpinns/trainers.py:626-677)
        u_test_losses = []
        start0, start1, report_time = time.time(), time.time(), 0.
        lossval = None
        for i, active_ in enumerate(scheduler):
            if active_ is not None:
                self.set_active(active_, i)
            if i == 0:
                u_test_losses, start1, report_time = self._report(
                    i, u_test_losses, start0, start1, report_time, self.u_exact, self.x_batch_test, lossval)
            lossval = self.step()
            u_test_losses, start1, report_time = self._report(
                i + 1, u_test_losses, start0, start1, report_time, self.u_exact, self.x_batch_test, lossval)
            if getattr(c, "save_models", False) and (i + 1) % c.model_save_freq == 0:        # off by default
                self.save_model(i + 1, self.inputs.active, u_test_losses)

        torch.cuda.synchronize()
        logger.info(f"[i: {c.n_steps}/{c.n_steps}] Training complete")
        self.u_test_losses = u_test_losses
        return self.export_all_params()

    # ---- checkpoints in the reference's format (fbpinns/trainers_base.py:64-69, trainers.py:721) ---------------------
    def _adam_trees(self):
        "mu / nu as trees with the structure of all_params['trainable']"
        trees = []
        for buf, pbuf in ((self.adam.mu, self.adam.pmu), (self.adam.nu, self.adam.pnu)):
            t = {"network": {"subdomain": {"layers": networks.from_kernel_layers(self.c.network, unpack_params(self.value_plan, buf))}}}
            prob = self.all_params["trainable"].get("problem", {})
            if prob:
                t["problem"], off = {}, 0
                for k, v in prob.items():
                    t["problem"][k] = pbuf[off:off + v.numel()].view(v.shape)
                    off += v.numel()
            trees.append(t)
        return trees

    def save_model(self, i, active, u_test_losses=(), path=None):
        "model_{i:08d}.jax = pickle of (i, all_params, optax-adam state, active, u_test_losses), numpy leaves"
        from .util import checkpoint
        mu, nu = self._adam_trees()
        path = path or os.path.join(self.c.model_out_dir, f"model_{i:08d}.jax")
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        checkpoint.save_model(path, i, self.export_all_params(), mu, nu, int(self.adam.count.item()), active, u_test_losses)
        return path

    def load_model(self, path):
        """Resume from a checkpoint written by save_model or by the reference (same network and decomposition): parameters,
        problem trainables and the Adam state go back into the packed device buffers.  Returns (i, active, u_test_losses);
        call set_active(active) afterwards."""
        from .util import checkpoint
        i, ap, (count, mu, nu), active, losses = checkpoint.load_model(path)
        dev = self.params.device
        tens = lambda tree: [tuple(torch.as_tensor(np.asarray(t), dtype=torch.float32, device=dev) for t in leaf)
                             for leaf in tree["network"]["subdomain"]["layers"]]

        def kernel_rows(tree):
            probe = {"static": self.all_params["static"], "trainable": {"network": {"subdomain": {"layers": tens(tree)}}}}
            return pack_params(self.value_plan, networks.kernel_layers(self.c.network, probe, dev)[2])
        self.params.copy_(kernel_rows(ap["trainable"]))
        self.adam.mu.copy_(kernel_rows(mu))
        self.adam.nu.copy_(kernel_rows(nu))
        if self.c.network.ACTIVATION == "fourier_tanh":      # the static feature rows carry no optimiser state
            n0 = self.layer_sizes[0] * self.layer_sizes[1] + self.layer_sizes[1]
            self.adam.mu[:, :n0] = 0
            self.adam.nu[:, :n0] = 0
        self.adam.count.fill_(int(count))
        if self.prob_flat is not None:
            for dst, tree in ((self.prob_flat, ap["trainable"]), (self.adam.pmu, mu), (self.adam.pnu, nu)):
                flat = torch.cat([torch.as_tensor(np.asarray(tree["problem"][k]), dtype=torch.float32).reshape(-1)
                                  for k in self.all_params["trainable"]["problem"]])
                dst[:flat.numel()].copy_(flat.to(dev))
        return i, active, losses

    def export_all_params(self):
        "all_params with the reference's pytree leaves: layers = [(w (m,out,in), b (m,out), activation parameters...), ...]"
        layers = networks.from_kernel_layers(self.c.network, unpack_params(self.value_plan, self.params))
        self.all_params["trainable"]["network"]["subdomain"]["layers"] = layers
        return self.all_params

    # ---- reporting / test (fbpinns/trainers.py:689-774, value-only twin of the hot path) ---------------------
    def _report(self, i, u_test_losses, start0, start1, report_time, u_exact, x_batch_test, lossval):
        c = self.c
        summary_, test_ = [(i % f == 0) for f in [c.summary_freq, c.test_freq]]
        if summary_ or test_:
            if i != 0 and summary_:
                torch.cuda.synchronize()
                rate = c.summary_freq / (time.time() - start1 - report_time)
                self._print_summary(i, float(lossval.item()), rate, start0)
                self.last_rate = rate
                start1, report_time = time.time(), 0.
            if test_:
                start2 = time.time()
                u_test = self.evaluate(x_batch_test)
                if u_exact is not None:
                    l1 = torch.mean(torch.abs(u_exact - u_test)).item()
                    l1n = l1 / u_exact.std(unbiased=False).item()          # jnp.std: population standard deviation
                    # row layout of the reference (fbpinns/trainers.py:757): [i, pstep, fstep, time, l1, l1n]; the
                    # per-step FLOP counters pstep / fstep are not tracked here
                    u_test_losses.append([i, 0, 0, time.time() - start0, l1, l1n])
                    logger.info(f"[i: {i}/{c.n_steps}] test l1: {l1:.5f} (normalised {l1n:.5f})")
                report_time += time.time() - start2
        return u_test_losses, start1, report_time

    def evaluate(self, x_batch):
        """Value-only FBPINN solution with ALL subdomains (FBPINN_model_jit / analysis.FBPINN_solution twin,
        fbpinns/trainers.py:314-320, 727-737): returns constrained u (n, ud)."""
        dd, plan = self.dd, self.value_plan
        # the takes are cached only for the very tensor object they were built for, unmodified since (identity + version
        # counter; the cache holds a reference, so the address cannot be reused) — e.g. the trainer's own x_batch_test
        cached = self._test_eval
        if cached is None or cached[0] is not x_batch or cached[1] != x_batch._version:
            x = x_batch.to(dd.device, torch.float32).contiguous()
            _, mc = dd.inside_count(x)
            _, a_ims, f_ims, all_ims, pos = active_set_algebra(np.ones(dd.m, dtype=int), mc.cpu().numpy())
            takes = DeviceTakes(dd, x, pos, all_ims, len(a_ims), tile_points=plan.tile_points)
            self._test_eval = (x_batch, x_batch._version, ConstraintEvaluator(plan, takes, x, dd, activation_cache=False), x)
        _, _, ev, x = self._test_eval
        with torch.no_grad():
            u = ev.forward(self.params)
            return self.c.problem.constraining_fn(self.all_params, x, u)
