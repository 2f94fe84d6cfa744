"""
ORACLE — TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH.

CPU restatement (numpy for index/integer work, torch-CPU for the floating-point model) of the
FBPINN training-step hot path of benmoseley/FBPINNs v0.2.0, written from the behaviour of the
reference (each function cites the reference file:line it follows; paths are relative to the
reference checkout).

Who may import this package: `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` — and there only as the checker / the CPU arm, never as the
thing measured as "ours" or shipped.  Nothing under `fbpinns_b200/` imports it.

Pinning status (see DESIGN.md §"Oracle"):
  * the reference itself cannot be imported in this image (it needs jax + optax, neither is
    installed and there is no network);
  * the oracle is pinned against the reference's own source executed here under a numpy-backed
    shim of the small jax subset it uses (`tests/golden/make_golden.py`, `make_golden_shim.py`):
    per-pair model values, window, norm/unnorm, `get_jmaps`, `_get_level_params`, the batched
    `inside_points / inside_models`, `get_inputs`, `_get_x_batch`, `_get_update_inputs` and the
    schedulers are the reference's code run verbatim; derivatives (ujs) and parameter gradients
    are pinned by central finite differences of those reference values in float64;
  * the only self-checking code in the reference (`fbpinns/decompositions_base.py:87-128`, pair
    ordering == row-major nonzero of the dense inside mask) is restated as a test;
  * optax's Adam and jax.random's threefry are third-party code absent from the reference tree:
    Adam is restated from its published formula ("parity unpinned" for Adam bit patterns), and
    parameter initialisation is always an explicit input (fbpinns_b200/util/jax_prng.py restates the threefry
    generator and is pinned by its known-answer vectors only).
"""
