"""CPU tests of the multi-GPU host logic with the gloo backend, world_size 2 (no GPU): ownership, halo lists and
the row exchange must reproduce the single-process row sums, and return owners' cotangents to the sharers."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_takes
from fbpinns_b200 import configs
from fbpinns_b200.parallel import Shard, build_halo_lists, HaloExchange, replicate_rows, peer_halo_layout


def _geometry(world):
    c = configs.cfg3_burgers(n_sub=(6, 4), n_pts=(40, 30))
    decomp = ref_takes.rectangular_init_params(**c.decomposition_init_kwargs)
    m = decomp["m"]
    xs = [np.linspace(-1, 1, 40), np.linspace(0, 1, 30)]
    x = np.stack(np.meshgrid(*xs, indexing="ij"), -1).reshape(-1, 2).astype(np.float32)
    mask = ref_takes.inside_mask(decomp, x, np.arange(m))               # (n, m)
    blocks = [Shard(j, world).block(m) for j in range(world)]
    inside = np.stack([mask[:, lo:hi].any(1) for lo, hi in blocks])   # (world, n)
    return x, mask, blocks, inside


def test_halo_lists_are_consistent():
    for world in (2, 3, 4):
        x, mask, blocks, inside = _geometry(world)
        halos = [build_halo_lists(inside, r) for r in range(world)]
        owner = halos[0]["owner"]
        assert (owner >= 0).all()
        # every point is owned exactly once, by the lowest rank that contains it
        owned = np.zeros(len(x), int)
        for r, h in enumerate(halos):
            owned[h["local_ips"][h["owned_local"]]] += 1
            assert (inside[:r, h["local_ips"][h["owned_local"]]] == False).all()
        assert (owned == 1).all()
        # send list of r towards j == recv list of j from r, as global point indices in the same order
        for r in range(world):
            for j in range(world):
                if j == r:
                    continue
                a = halos[r]["local_ips"][halos[r]["send"][j]]
                b = halos[j]["local_ips"][halos[j]["recv"][r]]
                assert np.array_equal(a, b)


def test_peer_halo_layout_simulated_exchange():
    """Index logic of the peer-memory halo exchange (fbp_halo_push / fbp_halo_pull): every rank's layout is built here and the
    two kernels are replayed in numpy (push = rows stored at the receiver's offsets, block by block; pull = CSR sums /
    copies): the owners must end with the full row sums, the sharers with the owners' values, and no two senders may touch
    the same receive-region row."""
    for world in (2, 3, 4):
        x, mask, blocks, inside = _geometry(world)
        halos = [build_halo_lists(inside, r) for r in range(world)]
        cnt = np.array([[0 if j == r else len(halos[r]["send"][j]) for j in range(world)] for r in range(world)])
        lays = [peer_halo_layout(halos[r], cnt, r, rows_per_block=37) for r in range(world)]
        V = 3
        pv = (np.arange(mask.shape[0])[:, None] * 0.001 + np.arange(mask.shape[1])[None, :] * 1.0 + 0.5)
        rows, total = [], []
        for r in range(world):
            lo, hi = blocks[r]
            lips = halos[r]["local_ips"]
            loc = (mask[:, lo:hi] * pv[:, lo:hi]).sum(1)[lips]
            rows.append(np.stack([loc, 2 * loc, -loc], 1))
            total.append((mask * pv).sum(1)[lips])
        region = [np.full((lays[r]["region_rows"], V), np.nan) for r in range(world)]
        written = [np.zeros(lays[r]["region_rows"], dtype=int) for r in range(world)]
        # forward push (every CTA of every rank), then forward pull
        for r in range(world):
            for j, p0, p1, nb in lays[r]["fwd_blocks"]:
                dst = lays[r]["rows_fwd_dst"][j] + np.arange(p0, p1)
                region[j][dst] = rows[r][lays[r]["fwd_send"][p0:p1]]
                written[j][dst] += 1
        for r in range(world):
            lay = lays[r]
            assert (written[r][:lay["my_bwd_off_rows"]] == 1).all() and (written[r][lay["my_bwd_off_rows"]:] == 0).all()
            for t, row in enumerate(lay["fwd_tgt"]):
                for s_ in range(lay["fwd_ptr"][t], lay["fwd_ptr"][t + 1]):
                    rows[r][row] += region[r][lay["fwd_pos"][s_]]
            own = halos[r]["owned_local"]
            assert np.allclose(rows[r][own, 0], total[r][own]) and np.allclose(rows[r][own, 2], -total[r][own])
        # reverse: owners hold g(point); push to the sharers' reverse regions, pull = copy
        back = []
        for r in range(world):
            lips, own = halos[r]["local_ips"], halos[r]["owned_local"]
            back.append(np.where(own, np.sin(lips * 0.37), 0.0)[:, None].repeat(V, 1))
        for r in range(world):
            for j, p0, p1, nb in lays[r]["bwd_blocks"]:
                dst = lays[r]["rows_bwd_dst"][j] + np.arange(p0, p1)
                region[j][dst] = back[r][lays[r]["bwd_send"][p0:p1]]
                written[j][dst] += 1
        for r in range(world):
            lay = lays[r]
            assert (written[r] == 1).all()
            for t, row in enumerate(lay["fwd_send"]):
                back[r][row] = region[r][lay["my_bwd_off_rows"] + lay["bwd_pos"][t]]
            assert np.allclose(back[r][:, 0], np.sin(halos[r]["local_ips"] * 0.37))
            assert lay["fwd_mask"] == sum(1 << j for j in range(world) if j != r and cnt[j][r]) and \
                lay["bwd_mask"] == sum(1 << j for j in range(world) if j != r and cnt[r][j])


def test_pair_count_balanced_blocks():
    "Shard.balance: contiguous blocks with (nearly) equal pair counts, every rank at least one subdomain, same on all ranks"
    rng = np.random.default_rng(0)
    for world in (2, 3, 4, 8):
        pairs = rng.integers(100, 3000, size=97)
        pairs[:12] //= 3                          # a light edge
        b = Shard(0, world).balance(pairs)
        assert b[0] == 0 and b[-1] == len(pairs) and (np.diff(b) >= 1).all()
        assert np.array_equal(b, Shard(world - 1, world).balance(pairs))
        loads = np.array([pairs[b[j]:b[j + 1]].sum() for j in range(world)])
        eq = np.array([pairs[(j * 97) // world:((j + 1) * 97) // world].sum() for j in range(world)])
        assert loads.max() <= eq.max() + pairs.max()
        assert loads.max() / loads.mean() < 1.0 + 1.5 * world * pairs.max() / pairs.sum()
        sh = Shard(1, world)
        sh.balance(pairs)
        assert sh.block(len(pairs)) == (int(b[1]), int(b[2]))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, mask, blocks, inside = _geometry(world)
        shard = Shard(rank, world)
        halo = build_halo_lists(inside, rank)
        ex = HaloExchange(halo, shard, torch.device("cpu"))
        lo, hi = blocks[rank]
        lips = halo["local_ips"]
        # per-pair value f(point, model); local partial row sums over this rank's models
        pv = (np.arange(mask.shape[0])[:, None] * 0.001 + np.arange(mask.shape[1])[None, :] * 1.0 + 0.5)
        local = (mask[:, lo:hi] * pv[:, lo:hi]).sum(1)[lips]
        total = (mask * pv).sum(1)[lips]
        rows = torch.tensor(np.stack([local, 2 * local], 1), dtype=torch.float64)
        ex.forward_add(rows)
        own = halo["owned_local"]
        ok_fwd = np.allclose(rows.numpy()[own, 0], total[own]) and np.allclose(rows.numpy()[own, 1], 2 * total[own])
        # reverse: owners hold g(point); sharers must receive it
        g = np.sin(lips * 0.37)
        back = torch.tensor(np.where(own, g, 0.0)[:, None].repeat(3, 1), dtype=torch.float64)
        ex.backward_return(back)
        ok_bwd = np.allclose(back.numpy()[:, 0], g)
        # "replicated" loss evaluation: owners' rows -> the full array on every rank; a scalar of the full array
        # differentiates back to exactly the owned rows (multi-constraint problems / problem trainables)
        n = mask.shape[0]
        owned_global = torch.as_tensor(lips[own], dtype=torch.long)
        vals = torch.tensor(np.stack([np.cos(lips[own] * 0.11), lips[own] * 1.0], 1), dtype=torch.float64, requires_grad=True)
        full = replicate_rows(vals, owned_global, n)
        wts = torch.arange(1, n + 1, dtype=torch.float64)[:, None] * torch.tensor([[1.0, -2.0]], dtype=torch.float64)
        (full * wts).sum().backward()
        ok_rep = (np.allclose(full.detach().numpy()[:, 0], np.cos(np.arange(n) * 0.11)) and
                  np.allclose(full.detach().numpy()[:, 1], np.arange(n) * 1.0) and
                  np.allclose(vals.grad.numpy(), wts.numpy()[lips[own]]))
        q.put((rank, bool(ok_fwd), bool(ok_bwd and ok_rep), int(own.sum()), int(len(lips))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_halo_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
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
    x, mask, blocks, inside = _geometry(world)
    assert sum(r[3] for r in res) == len(x)                  # every point owned once
    for rank, ok_fwd, ok_bwd, n_own, n_loc in res:
        assert ok_fwd and ok_bwd, (rank, ok_fwd, ok_bwd)
        assert n_loc > n_own or rank == 0
