"""
Host-side engine over the C ABI (include/fbpinn_b200.h): plans, device takes, the per-constraint evaluator
(forward jets + reverse pass as a torch.autograd.Function) and Adam on the packed parameter layout.

This is the Python mirror of the seam the reference exposes as
    FBPINN_forward(all_params, x_batch, takes, model_fns, jmaps) -> ujs        fbpinns/trainers.py:197-203
under value_and_grad (:292) and the optimiser lines (:294-295).  All arithmetic happens in libfbpinn_b200.so;
torch only owns the device buffers, the stream and the autograd tape that links the kernels to the user's
`Problem.loss_fn` / `constraining_fn`.
"""

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import FbpError, PlanDesc, TakesView, ptr, stream_ptr, check
from .jets import JetSpec

I32 = torch.int32


_DEVICE_INFO = {}


def device_info():
    "properties of torch's current CUDA device (cached per device ordinal: the query costs milliseconds)"
    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
    if dev not in _DEVICE_INFO:
        lib = _lib.load()
        sm, major, minor, smem = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        check(lib.fbp_device_info(C.byref(sm), C.byref(major), C.byref(minor), C.byref(smem)), "fbp_device_info")
        _DEVICE_INFO[dev] = dict(sm_count=sm.value, cc=(major.value, minor.value), smem_optin=smem.value)
    return _DEVICE_INFO[dev]


# --------------------------------------------------------------------------------------------------- plan

class Plan:
    """Network shape + jet spec (the static arguments `model_fns`/`jmaps` of the reference's FBPINN_forward)."""

    ACTIVATIONS = {"tanh": 0, "adaptive_tanh": 1, "sin": 2, "adaptive_sin": 3, "fourier_tanh": 4}

    def __init__(self, layer_sizes, jet: JetSpec, kernel="auto", activation="tanh"):
        """activation: the reference Network the packed rows belong to — "tanh" FCN, "adaptive_tanh" AdaptiveFCN, "sin"
        SIREN, "adaptive_sin" AdaptiveSIREN, "fourier_tanh" FourierFCN (layer_sizes then starts [xd, 2*n_features, ...]
        and layer 0 holds the static feature layer).  Everything but "tanh" runs on the generic kernel family."""
        lib = _lib.load()
        self.layer_sizes = [int(v) for v in layer_sizes]
        self.jet = jet
        self.activation = activation
        if len(self.layer_sizes) - 1 > _lib.FBP_MAX_LAYERS:
            raise FbpError(f"too many layers ({len(self.layer_sizes) - 1} > {_lib.FBP_MAX_LAYERS})")
        if jet.C > _lib.FBP_MAX_COMP:
            raise FbpError(f"too many jet components ({jet.C} > {_lib.FBP_MAX_COMP})")
        d = PlanDesc()
        d.xd, d.ud, d.n_layers = jet.xd, jet.ud, len(self.layer_sizes) - 1
        for i, v in enumerate(self.layer_sizes):
            d.layer_sizes[i] = v
        d.activation, d.window, d.n_comp = self.ACTIVATIONS[activation], 0, jet.C
        for c in range(jet.C):
            d.comp_k[c], d.comp_l[c] = jet.comp_k[c], jet.comp_l[c]
        h = C.c_void_p()
        check(lib.fbp_plan_create(C.byref(h), C.byref(d)), "fbp_plan_create")
        self._h = h
        self.P = int(lib.fbp_plan_param_count(h))
        self.n_extra = int(lib.fbp_plan_n_extra(h))
        self.set_kernel(kernel if activation == "tanh" else "auto")

    def set_kernel(self, kernel):
        lib = _lib.load()
        mode = {"auto": 0, "generic": 1, "tiled": 2, "tensor": 3, "tensor-full": 4}[kernel]
        self.has_tensor = bool(lib.fbp_plan_has_tensor(self._h))
        if mode in (3, 4) and not self.has_tensor:
            # "tensor" on a trainer means "where an instance exists" (e.g. not for a boundary constraint with other
            # jets); the family actually used is always visible as Plan.kernel
            kernel, mode = "auto", 0
        check(lib.fbp_plan_set_kernel(self._h, mode), "fbp_plan_set_kernel")
        self.kernel = kernel
        self.forward_family = ("generic", "tiled", "tensor")[int(lib.fbp_plan_forward_family(self._h))]
        self.reverse_family = ("generic", "tiled", "tensor")[int(lib.fbp_plan_reverse_family(self._h))]
        self.is_fast = bool(lib.fbp_plan_is_fast(self._h)) and mode != 1
        self.tile_points = int(lib.fbp_plan_tile_points(self._h))
        self.scratch_per_pair = int(lib.fbp_plan_scratch_per_pair(self._h))
        self.cache_per_pair = int(lib.fbp_plan_cache_per_pair(self._h))

    @property
    def handle(self):
        return self._h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().fbp_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


def pack_params(plan, layers):
    """[(w (m,out,in), b (m,out), [activation parameters (m,out) ...]), ...] CUDA float32 -> packed (m, P)."""
    lib = _lib.load()
    m = layers[0][0].shape[0]
    packed = torch.empty((m, plan.P), dtype=torch.float32, device=layers[0][0].device)
    ws = [leaf[0].contiguous().float() for leaf in layers]
    bs = [leaf[1].contiguous().float() for leaf in layers]
    n = len(layers)
    wp = (C.c_void_p * n)(*[w.data_ptr() for w in ws])
    bp = (C.c_void_p * n)(*[b.data_ptr() for b in bs])
    check(lib.fbp_pack_params(plan.handle, m, wp, bp, ptr(packed), stream_ptr()), "fbp_pack_params")
    for l, leaf in enumerate(layers):
        if len(leaf) != 2 + plan.n_extra:
            raise FbpError(f"layer {l} has {len(leaf)} leaves, the plan's activation expects {2 + plan.n_extra}")
        for e in range(plan.n_extra):
            v = leaf[2 + e].contiguous().float()
            check(lib.fbp_pack_extra(plan.handle, m, l, e, ptr(v), ptr(packed), 1, stream_ptr()), "fbp_pack_extra")
    return packed


def unpack_params(plan, packed):
    lib = _lib.load()
    m = packed.shape[0]
    ls = plan.layer_sizes
    ws = [torch.empty((m, o, i), dtype=torch.float32, device=packed.device) for i, o in zip(ls[:-1], ls[1:])]
    bs = [torch.empty((m, o), dtype=torch.float32, device=packed.device) for o in ls[1:]]
    n = len(ws)
    wp = (C.c_void_p * n)(*[w.data_ptr() for w in ws])
    bp = (C.c_void_p * n)(*[b.data_ptr() for b in bs])
    packed = packed.contiguous()
    check(lib.fbp_unpack_params(plan.handle, m, ptr(packed), wp, bp, stream_ptr()), "fbp_unpack_params")
    extras = [[] for _ in ws]
    for l, o in enumerate(ls[1:]):
        for e in range(plan.n_extra):
            v = torch.empty((m, o), dtype=torch.float32, device=packed.device)
            check(lib.fbp_pack_extra(plan.handle, m, l, e, ptr(v), ptr(packed), 0, stream_ptr()), "fbp_pack_extra")
            extras[l].append(v)
    return [(w, b) + tuple(ex) for w, b, ex in zip(ws, bs, extras)]


# --------------------------------------------------------------------------------------------------- decomposition on device

class DeviceDecomposition:
    """Static per-subdomain records for the kernels: [m][2*xd+3] = xmin, xmax, flag, unnorm mu, unnorm sd
    (float32 casts of RectangularDecompositionND's `params[0,1,4,5]`, fbpinns/decompositions.py:135-181)."""

    def __init__(self, params, pou, device):
        xmins, xmaxs, _, _, flags, unnorms = [torch.as_tensor(np.asarray(p.cpu() if torch.is_tensor(p) else p),
                                                              dtype=torch.float32) for p in params]
        self.m, self.xd = xmins.shape
        self.sub_static = torch.cat([xmins, xmaxs, flags, unnorms], dim=1).contiguous().to(device)
        pou_np = np.asarray(pou.cpu() if torch.is_tensor(pou) else pou).reshape(-1)
        self.pou_host = pou_np.astype(np.int64).astype(np.int32)
        if np.any(np.diff(self.pou_host) < 0):
            raise NotImplementedError("partition-of-unity ids must be non-decreasing in subdomain index "
                                      "(true for Rectangular/MultilevelRectangularDecompositionND)")
        self.pou = torch.as_tensor(self.pou_host, dtype=I32, device=device)
        self.npou = int(len(np.unique(self.pou_host)))
        self.device = device

    def inside_count(self, x, models=None, want_model_count=True):
        """(pt_count (n,), model_count (n_models,)) int32 device tensors."""
        lib = _lib.load()
        n = x.shape[0]
        nm = self.m if models is None else int(models.numel())
        pt = torch.empty(n, dtype=I32, device=self.device)
        mc = torch.empty(nm, dtype=I32, device=self.device) if want_model_count else None
        check(lib.fbp_inside_count(ptr(x), n, self.xd, ptr(self.sub_static), self.m, ptr(models), nm, ptr(pt),
                                   ptr(mc), stream_ptr()), "fbp_inside_count")
        return pt, mc


def nonzero_i32(count):
    lib = _lib.load()
    n = count.numel()
    out = torch.empty(n, dtype=I32, device=count.device)
    num = C.c_int64()
    check(lib.fbp_nonzero_i32(ptr(count), n, ptr(out), C.byref(num), stream_ptr()), "fbp_nonzero_i32")
    return out[:num.value].clone()


def gather_rows(src, idx):
    lib = _lib.load()
    src = src.contiguous()
    rf = int(np.prod(src.shape[1:])) if src.dim() > 1 else 1
    dst = torch.empty((idx.numel(),) + tuple(src.shape[1:]), dtype=torch.float32, device=src.device)
    check(lib.fbp_gather_rows(ptr(src), ptr(idx), idx.numel(), rf, ptr(dst), stream_ptr()), "fbp_gather_rows")
    return dst


# --------------------------------------------------------------------------------------------------- takes

def _split_tiles(ntiles, parts, split):
    """Tile counts of the `parts` items a subdomain of `ntiles` tiles is cut into.  "equal": equal parts (round 1 default).
    "guided": sizes fall off geometrically (weights 2^(n-1), ..., 2, 2 over n = parts + 1 items), so that with the
    longest-first launch order the big items fill the whole waves and the small ones the last, partial wave — the
    partial-wave tail of equal items is one candidate for what the reverse kernel loses at small per-GPU sizes.
    Experimental knob (FBP_ITEM_SPLIT=guided): a list-scheduling model of the N = 8 case predicts NO gain over the equal
    split with LPT order (packing efficiency 0.98 already), so it stays off until measured."""
    if split == "guided" and parts >= 2 and ntiles >= parts + 1:
        n = parts + 1
        w = np.array([2.0 ** (n - 1 - i) for i in range(n - 1)] + [2.0])
        sizes = np.maximum(1, np.floor(ntiles * w / w.sum()).astype(np.int64))
        i = 0
        while sizes.sum() < ntiles:              # hand the remainder out, largest first
            sizes[i % n] += 1
            i += 1
        i = 0
        while sizes.sum() > ntiles:              # only when the floor of 1 over-allocated
            if sizes[i % n] > 1:
                sizes[i % n] -= 1
            i += 1
        return [int(v) for v in sorted(sizes, reverse=True)]
    size = -(-ntiles // parts)
    out = []
    left = ntiles
    while left > 0:
        out.append(min(size, left))
        left -= out[-1]
    return out


def build_work_items(sub_off, m_active, tile_points, target_items, split=None):
    """Work list for the tiled kernels: (subdomain position, first pair, pair count, split index) rows, subdomain
    major.  A subdomain is cut into chunks of whole tiles so that the list has about `target_items` entries when the
    problem is small and one entry per subdomain when it is large; `split` ("equal" | "guided", default from
    FBP_ITEM_SPLIT, else "equal") chooses the chunk sizes (_split_tiles).  Also returns the launch orders (longest item
    first, LPT) for the forward (all items) and reverse (active items) kernels."""
    if split is None:
        split = os.environ.get("FBP_ITEM_SPLIT", "equal")
    sub_off = np.asarray(sub_off, dtype=np.int64)
    m_all = len(sub_off) - 1
    s = int(sub_off[-1])
    chunk0 = max(1, -(-s // max(1, target_items)))
    chunk0 = -(-chunk0 // tile_points) * tile_points
    cnts = np.diff(sub_off)
    parts_all = np.minimum(-(-cnts // chunk0), -(-cnts // tile_points))
    if split == "equal" and (parts_all <= 1).all():
        # common large-problem case, vectorised: one work item per non-empty subdomain
        nz = np.nonzero(cnts > 0)[0]
        items = np.stack([nz, sub_off[:-1][nz], cnts[nz], np.zeros_like(nz)], axis=1)
        sub_item_off = np.concatenate([[0], np.cumsum(cnts > 0)])
    else:
        items, sub_item_off = [], [0]
        for sp in range(m_all):
            a, b = int(sub_off[sp]), int(sub_off[sp + 1])
            cnt = b - a
            if cnt > 0:
                parts = -(-cnt // chunk0)
                ntiles = -(-cnt // tile_points)
                for k, nt in enumerate(_split_tiles(ntiles, min(parts, ntiles), split)):
                    c = min(nt * tile_points, b - a)
                    items.append((sp, a, c, k))
                    a += c
            sub_item_off.append(len(items))
    items = np.asarray(items, dtype=np.int32).reshape(-1, 4)
    sub_item_off = np.asarray(sub_item_off, dtype=np.int32)
    n_items_active = int(sub_item_off[m_active])
    order_fwd = np.argsort(-items[:, 2], kind="stable").astype(np.int32)
    order_bwd = np.argsort(-items[:n_items_active, 2], kind="stable").astype(np.int32)
    return items, sub_item_off, n_items_active, order_fwd, order_bwd


def launch_records(items, order, sub_ids):
    """Launch records of the tensor kernels (include/fbpinn_b200.h d_launch_fwd / d_launch_bwd): row b = what block b of
    the launch needs — (first pair, pair count, global subdomain index, item) of work item order[b] — so that the kernel
    reaches its parameters with two dependent loads instead of four (launch order -> work list -> subdomain index -> row)."""
    items = np.asarray(items, dtype=np.int32).reshape(-1, 4)
    order = np.asarray(order, dtype=np.int32)
    it = items[order]
    ims = np.asarray(sub_ids, dtype=np.int32)
    return np.stack([it[:, 1], it[:, 2], ims[it[:, 0]], order], axis=1).astype(np.int32).reshape(-1, 4)


def one_item_per_subdomain(sub_item_off, m_active, n_items_active):
    """True when item i IS active subdomain position i (what FBP_BWD_DIRECT asserts to the library)."""
    return bool(n_items_active == m_active and
                np.array_equal(np.asarray(sub_item_off)[:m_active + 1], np.arange(m_active + 1)))


class DeviceTakes:
    """One constraint's takes on the device: the reference-order arrays (m_take, n_take, p_take, np_take — bit-exact
    with fbpinns/trainers.py:336-391 after the per-constraint split of :544-571) and the subdomain-sorted view."""

    def __init__(self, decomp: DeviceDecomposition, x, pos_of_model, all_ims, m_active, tile_points=128,
                 target_items=None):
        lib = _lib.load()
        dev = decomp.device
        self.n = int(x.shape[0])
        self.m_all, self.m_active, self.npou = int(len(all_ims)), int(m_active), decomp.npou
        self.sub_ids_host = np.asarray(all_ims, dtype=np.int32)
        self.sub_ids = torch.as_tensor(self.sub_ids_host, dtype=I32, device=dev)
        pos_d = torch.as_tensor(np.asarray(pos_of_model, dtype=np.int32), dtype=I32, device=dev)
        b = C.c_void_p()
        s, q = C.c_int64(), C.c_int64()
        try:
            check(lib.fbp_takes_begin(C.byref(b), ptr(x), self.n, decomp.xd, ptr(decomp.sub_static), decomp.m,
                                      ptr(pos_d), ptr(decomp.pou), self.m_all, stream_ptr(), C.byref(s), C.byref(q)),
                  "fbp_takes_begin")
            self.s, self.q = int(s.value), int(q.value)
            mk = lambda k: torch.empty(max(k, 1), dtype=I32, device=dev)[:k]
            self.m_take, self.n_take, self.p_take = mk(self.s), mk(self.s), mk(self.s)
            self.np_take = mk(self.q)
            self.row_off, self.pt_row_off, self.sub_off = mk(self.q + 1), mk(self.n + 1), mk(self.m_all + 1)
            self.spair_point, self.spair_row, self.spair_sub, self.pos = mk(self.s), mk(self.s), mk(self.s), mk(self.s)
            check(lib.fbp_takes_emit(b, ptr(self.m_take), ptr(self.n_take), ptr(self.p_take), ptr(self.np_take),
                                     ptr(self.row_off), ptr(self.pt_row_off), ptr(self.sub_off),
                                     ptr(self.spair_point), ptr(self.spair_row), ptr(self.spair_sub), ptr(self.pos),
                                     stream_ptr()), "fbp_takes_emit")
        finally:
            if b:
                lib.fbp_takes_destroy(b)
        self.sub_off_host = self.sub_off.cpu().numpy()
        self.s_active = int(self.sub_off_host[self.m_active])
        self.set_tiling(tile_points, target_items)

    def set_tiling(self, tile_points, target_items=None):
        if target_items is None:
            # FBP_TARGET_ITEMS: tuning knob for A/B runs (work items the list should have at least, if possible)
            # 4 items per SM: measured on one rank's share of the 8-GPU run (512 subdomains): 738 us for the two subdomain
            # kernels against 825 us with 8 per SM (every extra item pays ~5-13 us of prologue / epilogue)
            target_items = int(os.environ.get("FBP_TARGET_ITEMS", 0)) or 4 * device_info()["sm_count"]
        items, sub_item_off, nia, order_fwd, order_bwd = build_work_items(self.sub_off_host, self.m_active, tile_points,
                                                                          target_items)
        dev = self.sub_ids.device
        self.items_host = items
        self.items = torch.as_tensor(items.reshape(-1), dtype=I32, device=dev)
        self.sub_item_off = torch.as_tensor(sub_item_off, dtype=I32, device=dev)
        self.item_order_fwd = torch.as_tensor(order_fwd, dtype=I32, device=dev)
        self.item_order_bwd = torch.as_tensor(order_bwd, dtype=I32, device=dev)
        self.n_items, self.n_items_active = int(items.shape[0]), nia
        # exactly one work item per active subdomain (the large-problem case): the reverse kernels can then write the gradient
        # rows themselves (FBP_BWD_DIRECT)
        self.one_item_per_sub = one_item_per_subdomain(sub_item_off, self.m_active, nia)
        self.launch_fwd = torch.as_tensor(launch_records(items, order_fwd, self.sub_ids_host).reshape(-1), dtype=I32, device=dev)
        self.launch_bwd = torch.as_tensor(launch_records(items, order_bwd, self.sub_ids_host).reshape(-1), dtype=I32, device=dev)
        self.tile_points = tile_points
        self._view = None

    def view(self):
        if self._view is None:
            v = TakesView()
            v.n, v.s, v.q, v.s_active = self.n, self.s, self.q, self.s_active
            v.m_all, v.m_active, v.npou = self.m_all, self.m_active, self.npou
            for name, t in [("d_m_take", self.m_take), ("d_np_take", self.np_take), ("d_sub_ids", self.sub_ids),
                            ("d_sub_off", self.sub_off), ("d_spair_point", self.spair_point),
                            ("d_spair_row", self.spair_row), ("d_spair_sub", self.spair_sub), ("d_pos", self.pos),
                            ("d_row_off", self.row_off), ("d_pt_row_off", self.pt_row_off), ("d_items", self.items),
                            ("d_sub_item_off", self.sub_item_off), ("d_item_order_fwd", self.item_order_fwd),
                            ("d_item_order_bwd", self.item_order_bwd), ("d_launch_fwd", self.launch_fwd),
                            ("d_launch_bwd", self.launch_bwd)]:
                setattr(v, name, t.data_ptr() if t.numel() else None)
            v.n_items, v.n_items_active = self.n_items, self.n_items_active
            self._view = v
        return self._view

    def reference_arrays(self):
        """(m_take, n_take, p_take, np_take, npou) as numpy int32 — the reference's `takes` tuple."""
        return (self.m_take.cpu().numpy(), self.n_take.cpu().numpy(), self.p_take.cpu().numpy(),
                self.np_take.cpu().numpy(), self.npou)


# --------------------------------------------------------------------------------------------------- per-constraint evaluator

class ConstraintEvaluator:
    """Everything the kernels need for one constraint between two active-set changes."""

    GENERIC_SCRATCH_FLOATS = 64 * 1024 * 1024   # 256 MB cap for the generic family's per-pair scratch

    def __init__(self, plan: Plan, takes: DeviceTakes, x, decomp: DeviceDecomposition, activation_cache=True):
        lib = _lib.load()
        self.plan, self.takes, self.decomp = plan, takes, decomp
        self.x = x.contiguous()
        dev = x.device
        V = plan.jet.C * plan.jet.ud
        self.V = V
        if takes.tile_points != plan.tile_points:
            takes.set_tiling(plan.tile_points)
        f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        self.dsum = f(max(takes.q, 1), plan.jet.C)
        self.pair_out = f(max(takes.s, 1), V)
        self.grow = f(max(takes.q, 1), V)
        nws = int(lib.fbp_backward_workspace_floats(plan.handle, C.byref(takes.view())))
        if nws < 0:
            raise FbpError("fbp_backward_workspace_floats failed")
        self.gpart = f(max(nws, 1))
        spp = plan.scratch_per_pair
        ns = min(max(takes.s, 128) * spp, max(self.GENERIC_SCRATCH_FLOATS, 128 * spp)) if spp else 0
        self.scratch = f(max(ns, 1))
        self.scratch_floats = ns
        # optional activation cache (tiled plans with two hidden layers): the forward kernel saves the last hidden
        # layer's jets, the reverse kernel TMA-loads them instead of recomputing the hidden GEMM
        ncache = takes.s * plan.cache_per_pair if activation_cache else 0
        self.cache = torch.zeros(ncache, dtype=torch.float32, device=dev) if ncache else None
        self.affine = None      # optional (n, 2C): jets of A and B of an affine constraining operator (set_affine)
        check(lib.fbp_window_sums(plan.handle, C.byref(takes.view()), ptr(self.x), ptr(decomp.sub_static),
                                  ptr(self.dsum), stream_ptr()), "fbp_window_sums")

    def set_affine(self, aff):
        "fuse the constraining operator A(x) u + B(x) (jets.AffineConstraining) into the reduce kernels"
        self.affine = None if aff is None else torch.cat([aff.Aj, aff.Bj], dim=1).contiguous().float()

    def forward(self, params):
        """params: packed (m, P) -> ujets (n, C*ud) of the subdomain sum (of the CONSTRAINED solution when an affine
        constraining operator has been attached with set_affine)."""
        lib = _lib.load()
        tv = self.takes.view()
        check(lib.fbp_forward(self.plan.handle, C.byref(tv), ptr(self.x), ptr(params), ptr(self.decomp.sub_static),
                              ptr(self.pair_out), ptr(self.scratch), self.scratch_floats, ptr(self.cache), stream_ptr()),
              "fbp_forward")
        ujets = torch.empty((self.takes.n, self.V), dtype=torch.float32, device=self.x.device)
        check(lib.fbp_reduce_forward(self.plan.handle, C.byref(tv), ptr(self.pair_out), ptr(self.dsum), ptr(self.affine),
                                     ptr(ujets), stream_ptr()), "fbp_reduce_forward")
        return ujets

    # set by the step object when this evaluator is the ONLY contribution to the gradient buffer and the work list has one
    # item per active subdomain: the reverse kernel writes the rows of `grads` itself (no zero fill, no reduction pass)
    direct_grads = False

    def supports_direct_grads(self):
        return bool(self.takes.one_item_per_sub and self.plan.reverse_family != "generic")

    def backward(self, ujets_bar, params, grads, accumulate=True):
        """Cotangent of ujets -> adds (or writes) the gradients of the active subdomains into grads (m_active, P)."""
        lib = _lib.load()
        flags = 1 if accumulate else 0
        if self.direct_grads:
            if not self.supports_direct_grads():
                raise FbpError("direct_grads set on an evaluator whose work list has not one item per active subdomain")
            flags = 8                                   # FBP_BWD_DIRECT
        tv = self.takes.view()
        ub = ujets_bar.contiguous().float()
        check(lib.fbp_reduce_backward(self.plan.handle, C.byref(tv), ptr(ub), ptr(self.dsum), ptr(self.affine),
                                      ptr(self.grow), stream_ptr()), "fbp_reduce_backward")
        check(lib.fbp_backward(self.plan.handle, C.byref(tv), ptr(self.x), ptr(params), ptr(self.decomp.sub_static),
                               ptr(self.grow), ptr(grads), flags, ptr(self.gpart), ptr(self.scratch),
                               self.scratch_floats, ptr(self.cache), stream_ptr()), "fbp_backward")

    def pair_values_reference_order(self):
        """Per-pair numerator jets in the reference's (point-sorted) pair order, after a forward()."""
        return self.pair_out[self.takes.pos.long()]


class _SubdomainSum(torch.autograd.Function):
    """ujets = kernels(params); reverse pass writes into `grads` (the packed gradient buffer of the active set)."""

    @staticmethod
    def forward(ctx, tape_hook, ev, params, grads):
        ctx.ev, ctx.params, ctx.grads = ev, params, grads
        return ev.forward(params)

    @staticmethod
    def backward(ctx, ujets_bar):
        ctx.ev.backward(ujets_bar, ctx.params, ctx.grads, accumulate=True)
        return None, None, None, None


def subdomain_sum(ev, params, grads, tape_hook):
    """Differentiable (w.r.t. the packed parameters, through `grads`) evaluation of one constraint.
    `tape_hook` is any scalar tensor with requires_grad=True: it only makes autograd schedule the reverse kernel."""
    return _SubdomainSum.apply(tape_hook, ev, params, grads)


# --------------------------------------------------------------------------------------------------- Adam

class PackedAdam:
    """optax.adam state on the packed layout: mu/nu (m, P) for the subdomain networks, (K,) for the problem's own
    trainables, ONE shared int32 step counter (the reference's global `count`, fbpinns/trainers.py:51-60)."""

    def __init__(self, m, P, n_problem, device, learning_rate=1e-3, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=device)
        self.mu, self.nu = z(m, P), z(m, P)
        self.pmu, self.pnu = z(max(n_problem, 1)), z(max(n_problem, 1))
        self.count = torch.zeros(1, dtype=I32, device=device)
        self.hyper = (float(learning_rate), float(b1), float(b2), float(eps), float(eps_root))

    def step(self, params, grads, active_ims_dev, prob_flat=None, prob_grad=None):
        lib = _lib.load()
        lr, b1, b2, eps, er = self.hyper
        if prob_flat is not None and prob_flat.numel() > 0:
            check(lib.fbp_adam_step(ptr(prob_flat), ptr(self.pmu), ptr(self.pnu), ptr(prob_grad), None, 1,
                                    prob_flat.numel(), ptr(self.count), 0, lr, b1, b2, eps, er, stream_ptr()),
                  "fbp_adam_step(problem)")
        check(lib.fbp_adam_step(ptr(params), ptr(self.mu), ptr(self.nu), ptr(grads), ptr(active_ims_dev),
                                active_ims_dev.numel(), params.shape[1], ptr(self.count), 1, lr, b1, b2, eps, er,
                                stream_ptr()), "fbp_adam_step")


def fma_peak_tflops(iters=20000, packed=False):
    "FP32 FMA micro-benchmark: scalar FFMA, or the packed FFMA2 (fma.rn.f32x2) when packed=True"
    lib = _lib.load()
    v = C.c_float()
    fn = lib.fbp_ffma2_peak if packed else lib.fbp_fma_peak
    check(fn(iters, C.byref(v), stream_ptr()), "fbp_fma_peak")
    return float(v.value)
