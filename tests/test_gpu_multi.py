"""Multi-GPU parity (needs >= 2 GPUs, otherwise skipped): the subdomain-sharded step (NCCL halo exchange) must
reproduce the single-GPU loss curve and parameters, and its first step the float64 ORACLE's loss and gradients (1e-5)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, name, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if name.endswith("-nccl"):          # the collective transport instead of the peer-memory one (fbp_halo_*)
        os.environ["FBP_HALO"] = "nccl"
        name = name[:-5]
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from fbpinns_b200 import configs
        from fbpinns_b200.trainers import FBPINNTrainer
        from fbpinns_b200.parallel import shard_trainer
        from fbpinns_b200.util.logger import logger
        logger.setLevel("WARNING")
        n_steps = 12

        def make(graph):
            if name == "cfg5":
                return configs.cfg5_poisson(n_sub=(8, 6), n_pts=(96, 72), device=f"cuda:{rank}", use_cuda_graph=graph)
            if name == "cfg2":
                return configs.cfg2_harmonic_oscillator_inverse(n_sub=10, n_pts=120, device=f"cuda:{rank}", use_cuda_graph=graph)
            if name == "cfg1":      # two constraints; the boundary point x = 0 is owned by rank 0 only: the others own none of it
                return configs.cfg1_harmonic_oscillator(n_sub=9, n_pts=90, device=f"cuda:{rank}", use_cuda_graph=graph)
            if name == "multilevel":        # two partitions of unity: rows are (point, level), exchanged in the dense layout
                from fbpinns_b200 import decompositions
                from fbpinns_b200.constants import get_subdomain_ws
                xs1 = [np.linspace(-1, 1, 3), np.linspace(0, 1, 2)]
                xs2 = [np.linspace(-1, 1, 6), np.linspace(0, 1, 4)]
                c = configs.cfg3_burgers(n_sub=(3, 2), n_pts=(40, 30), line_scheduler=False, device=f"cuda:{rank}",
                                         use_cuda_graph=graph)
                c.decomposition = decompositions.MultilevelRectangularDecompositionND
                c.decomposition_init_kwargs = dict(subdomain_xss=[xs1, xs2], subdomain_wss=[get_subdomain_ws(xs1, 2.9),
                                                                                          get_subdomain_ws(xs2, 2.9)], unnorm=(0., 3.))
                return c
            return configs.cfg3_burgers(n_sub=(6, 5), n_pts=(48, 40), line_scheduler=False, device=f"cuda:{rank}",
                                        use_cuda_graph=graph)
        # single-GPU reference trajectory (every rank computes the same one)
        ref = FBPINNTrainer(make(False)).setup()
        ref.set_active(np.ones(ref.dd.m, dtype=int))
        # float64 oracle on the trainer's own initial parameters: loss and gradients of the first step
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import common
        from fbpinns_b200.engine import unpack_params
        k = common.make_case(make(False), seed=0, multilevel=(name == "multilevel"))
        k.layers = [(w.cpu().numpy(), b.cpu().numpy()) for w, b in unpack_params(ref.value_plan, ref.params)]
        if ref.prob_flat is not None:
            k.prob_trainable = {kk: ref.prob_flat.detach().cpu().numpy()[i].astype(np.float32)
                                for i, kk in enumerate(ref.all_params["trainable"]["problem"])}
        o_loss, o_grads, o_gprob = common.oracle_loss_and_grads(k, torch.float64)
        ref_losses = [float(ref.step().item()) for _ in range(n_steps)]
        out = {}
        for graph in (False, True):
            tr = shard_trainer(FBPINNTrainer(make(graph)), rank, world).setup()
            tr.set_active(np.ones(tr.dd.m, dtype=int))
            losses = [float(tr.step().item())]
            lo, hi = tr.shard.block(tr.dd.m)
            g_err = 0.0
            if not graph:       # gradients of the first step (this rank's block of subdomains) against the oracle
                got = unpack_params(tr.value_plan, tr.update.grads[:len(tr.inputs.active_ims)].contiguous())
                rows = np.asarray(tr.inputs.active_ims)
                for (gw, gb), (rw, rb) in zip(got, o_grads):
                    g_err = max(g_err, float(np.max(np.abs(gw.cpu().numpy() - rw[rows])) / np.max(np.abs(rw))),
                                float(np.max(np.abs(gb.cpu().numpy() - rb[rows])) / np.max(np.abs(rb))))
            losses += [float(tr.step().item()) for _ in range(n_steps - 1)]
            perr = float((tr.params[lo:hi] - ref.params[lo:hi]).abs().max() / ref.params.abs().max())
            pe = 0.0
            if tr.prob_flat is not None:
                pe = float((tr.prob_flat - ref.prob_flat).abs().max())
            out[graph] = (losses, perr, pe, tr.update.graph is not None, g_err)
            halo0 = tr.inputs.evaluators[0].halo
            assert type(halo0).__name__ == ("HaloExchange" if os.environ.get("FBP_HALO") == "nccl" else "PeerHaloExchange")
        q.put((rank, ref_losses, out, o_loss))
        dist.barrier()
        torch.cuda.synchronize()
    except BaseException:
        import traceback
        traceback.print_exc()
        os._exit(1)
    import time
    time.sleep(1.0)              # let the queue feeder thread flush
    os._exit(0)                  # NCCL teardown with captured graphs alive can hang (see bench.py)


@pytest.mark.parametrize("name", ["cfg5", "cfg3", "cfg2", "cfg1", "cfg5-nccl", "multilevel"])
def test_sharded_step_matches_single_gpu(name):
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if world < 4 else 4
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue as _queue
    import time as _time
    res, t_end = [], _time.time() + 420
    while len(res) < world and _time.time() < t_end:
        try:
            res.append(q.get(timeout=2))
        except _queue.Empty:
            dead = [p for p in procs if p.exitcode not in (None, 0)]
            if dead:                                   # a worker crashed: do not wait for the others
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                raise AssertionError(f"worker exited with code {dead[0].exitcode}")
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.terminate()
    assert len(res) == world, "workers did not finish in time"
    for p in procs:
        assert p.exitcode == 0
    for rank, ref_losses, out, o_loss in res:
        for graph, (losses, perr, pe, captured, g_err) in out.items():
            rel = np.max(np.abs(np.array(losses) - np.array(ref_losses)) / np.abs(np.array(ref_losses)))
            # multi-constraint problems and problem trainables are evaluated on the replicated ujs (parallel.py): the
            # sharded loss is the global one for every problem
            assert rel < 1e-4, (name, rank, graph, rel, losses, ref_losses)
            # first step against the float64 oracle: loss and this rank's gradients within 1e-5
            assert abs(losses[0] - o_loss) <= 1e-5 * abs(o_loss), (name, rank, graph, losses[0], o_loss)
            assert g_err < 1e-5, (name, rank, graph, g_err)
            assert perr < 1e-4, (name, rank, graph, perr)
            assert pe < 1e-4, (name, rank, graph, pe)
            if graph:
                assert captured
