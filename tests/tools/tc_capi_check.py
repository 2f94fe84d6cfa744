"""Torch-free check of the tensor (tcgen05) kernels through the C ABI: a few independent subdomains with full tiles,
partial tails and a tiny one; forward outputs and reverse-pass gradients of every tensor variant against the tiled
(FFMA2) kernels on the same inputs.  Only needs numpy, libcudart and libfbpinn_b200.so (starts in ~1 s on a fresh box).

    python tests/tools/tc_capi_check.py            -> JSON lines (also appended to gpurun_out/tc_capi_check.jsonl)
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbpinns_b200 import _lib                               # noqa: E402  (no torch import on this path)
from fbpinns_b200._lib import PlanDesc, TakesView           # noqa: E402

rt = None
for name in ("libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
    try:
        rt = C.CDLL(name)
        break
    except OSError:
        continue
if rt is None:
    raise SystemExit("libcudart not found")
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
rt.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
rt.cudaGetErrorString.restype = C.c_char_p


def ck(rc, what):
    if rc != 0:
        raise SystemExit(f"{what}: cuda error {rc} {rt.cudaGetErrorString(rc).decode()}")


class Dev:
    "device copy of a numpy array"

    def __init__(self, arr):
        self.arr = np.ascontiguousarray(arr)
        self.p = C.c_void_p()
        ck(rt.cudaMalloc(C.byref(self.p), max(self.arr.nbytes, 16)), "cudaMalloc")
        ck(rt.cudaMemcpy(self.p, self.arr.ctypes.data_as(C.c_void_p), self.arr.nbytes, 1), "h2d")

    def get(self):
        out = np.empty_like(self.arr)
        ck(rt.cudaMemcpy(out.ctypes.data_as(C.c_void_p), self.p, out.nbytes, 2), "d2h")
        return out

    def fill_nan(self):
        ck(rt.cudaMemset(self.p, 0xff, self.arr.nbytes), "memset")


def rel(a, b):
    return float(np.nanmax(np.abs(a - b)) / max(np.abs(b).max(), 1e-30))


SPECS = {   # name -> (xd, jet components as (k, l) pairs in JetSpec order)
    "cfg5": (2, [(-1, -1), (0, -1), (1, -1), (0, 0), (1, 1)]),
    "cfg2": (1, [(-1, -1), (0, -1), (0, 0)]),
    "burgers": (2, [(-1, -1), (0, -1), (1, -1), (0, 0)]),
    "value": (2, [(-1, -1)]),
    "first": (2, [(-1, -1), (0, -1)]),
}


def main(spec="cfg5"):
    lib = _lib.load()
    rng = np.random.default_rng(0)
    xd, comps = SPECS[spec]
    H, Cj = 32, len(comps)
    P = H * xd + H + H * H + H + H + 1
    counts = [300, 128, 77, 1, 515]
    m, s = len(counts), sum(counts)
    # decomposition records [xmin(2), xmax(2), flag, un_mu, un_sd] and points strictly inside their boxes
    lo = rng.uniform(-1, 0, (m, xd)).astype(np.float32)
    hi = (lo + rng.uniform(0.5, 1.5, (m, xd))).astype(np.float32)
    sub_static = np.concatenate([lo, hi, np.ones((m, 1)), np.full((m, 1), 0.1), np.full((m, 1), 1.3)], axis=1).astype(np.float32)
    x = np.concatenate([lo[i] + (hi[i] - lo[i]) * rng.uniform(0.02, 0.98, (c, xd)) for i, c in enumerate(counts)]).astype(np.float32)
    params = np.zeros((m, P), dtype=np.float32)
    off = 0
    for fan_in, nel in ((xd, H * xd), (xd, H), (H, H * H), (H, H), (H, H), (H, 1)):
        params[:, off:off + nel] = rng.uniform(-1, 1, (m, nel)) / np.sqrt(fan_in)
        off += nel
    first = np.cumsum([0] + counts[:-1])
    items = np.array([[i, first[i], counts[i], 0] for i in range(m)], dtype=np.int32)
    d = dict(x=Dev(x), params=Dev(params), ss=Dev(sub_static), sub_ids=Dev(np.arange(m, dtype=np.int32)),
             spair=Dev(np.arange(s, dtype=np.int32)), items=Dev(items), sio=Dev(np.arange(m + 1, dtype=np.int32)),
             pair_out=Dev(np.zeros((s, Cj), np.float32)), grow=Dev(rng.standard_normal((s, Cj)).astype(np.float32)),
             gpart=Dev(np.zeros((m, P), np.float32)), grads=Dev(np.zeros((m, P), np.float32)),
             cache=Dev(np.zeros((s, H * Cj), np.float32)))
    tv = TakesView()
    tv.n, tv.s, tv.q, tv.s_active = s, s, s, s
    tv.m_all, tv.m_active, tv.npou = m, m, 1
    tv.d_sub_ids, tv.d_spair_point, tv.d_spair_row = d["sub_ids"].p, d["spair"].p, d["spair"].p
    tv.d_items, tv.d_sub_item_off = d["items"].p, d["sio"].p
    tv.d_item_order_fwd = tv.d_item_order_bwd = None
    tv.n_items, tv.n_items_active = m, m

    pd = PlanDesc()
    pd.xd, pd.ud, pd.n_layers = xd, 1, 3
    for i, v in enumerate([xd, H, H, 1]):
        pd.layer_sizes[i] = v
    pd.activation, pd.window, pd.n_comp = 0, 0, Cj
    for c, (k, l) in enumerate(comps):
        pd.comp_k[c], pd.comp_l[c] = k, l
    plan = C.c_void_p()
    _lib.check(lib.fbp_plan_create(C.byref(plan), C.byref(pd)), "fbp_plan_create")

    lines = []

    def emit(**kw):
        line = json.dumps(dict(spec=spec, **kw))
        print(line, flush=True)
        lines.append(line)
        outdir = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(outdir):
            with open(os.path.join(outdir, "tc_capi_check.jsonl"), "a") as f:
                f.write(line + "\n")

    def forward(mode, fwd_variant=1, cache=True):
        os.environ["FBP_TC_FWD"] = str(fwd_variant)
        _lib.check(lib.fbp_plan_set_kernel(plan, mode), "fbp_plan_set_kernel")
        d["pair_out"].fill_nan()
        d["cache"].fill_nan()
        _lib.check(lib.fbp_forward(plan, C.byref(tv), d["x"].p, d["params"].p, d["ss"].p, d["pair_out"].p, None, 0,
                                   d["cache"].p if cache else None, None), "fbp_forward")
        ck(rt.cudaDeviceSynchronize(), "sync after forward")
        return d["pair_out"].get(), d["cache"].get()

    def backward(mode, cache):
        _lib.check(lib.fbp_plan_set_kernel(plan, mode), "fbp_plan_set_kernel")
        d["grads"].fill_nan()
        _lib.check(lib.fbp_backward(plan, C.byref(tv), d["x"].p, d["params"].p, d["ss"].p, d["grow"].p, d["grads"].p, 0,
                                    d["gpart"].p, None, 0, d["cache"].p if cache else None, None), "fbp_backward")
        ck(rt.cudaDeviceSynchronize(), "sync after backward")
        return d["grads"].get()

    ref_out, ref_cache = forward(2)
    emit(step="forward tiled", nan=int(np.isnan(ref_out).sum()))
    out, cache = forward(3, 1)
    emit(step="forward tensor v1", rel=rel(out, ref_out), cache_rel=rel(cache, ref_cache), nan=int(np.isnan(out).sum()))
    g_ref = backward(2, cache=False)
    emit(step="backward tiled (recompute)", nan=int(np.isnan(g_ref).sum()))
    out, cache = forward(3, 2)
    emit(step="forward tensor v2 (pipelined)", rel=rel(out, ref_out), cache_rel=rel(cache, ref_cache), nan=int(np.isnan(out).sum()))
    g = backward(4, cache=False)
    seg = {"W0": (0, H * xd), "b0": (H * xd, H * xd + H), "W1": (H * xd + H, H * xd + H + H * H),
           "b1": (H * xd + H + H * H, H * xd + 2 * H + H * H), "W2": (H * xd + 2 * H + H * H, P - 1), "b2": (P - 1, P)}
    emit(step="backward tensor-full", rel=rel(g, g_ref), nan=int(np.isnan(g).sum()),
         **{k: rel(g[:, a:b], g_ref[:, a:b]) for k, (a, b) in seg.items()})
    return 0


if __name__ == "__main__":
    for name in (sys.argv[1:] or ["cfg5"]):
        if name == "all":
            for n in SPECS:
                main(n)
        else:
            main(name)
