"""Stand-alone driver of fbp_tc_selftest (the tcgen05 MMA form of csrc/fbp_tc.cuh) that needs neither torch nor
pytest: ctypes on libcudart + libfbpinn_b200.so, numpy for the float64 reference.  Each variant runs in its own
process because a failed mbarrier wait traps (by design) and takes the CUDA context with it.

    python tests/tools/tc_selftest.py [variant ...]        -> one JSON line per variant (also appended to
                                                              gpurun_out/tc_selftest.jsonl when that directory exists)
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run_variant(variant, mn=False):
    rt = None
    for name in ("libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = C.CDLL(name)
            break
        except OSError:
            continue
    if rt is None:
        raise SystemExit("libcudart not found")
    lib = C.CDLL(os.path.join(ROOT, "fbpinns_b200", "csrc", "libfbpinn_b200.so"))
    fn = lib.fbp_tc_selftest_g if mn else lib.fbp_tc_selftest
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.fbp_last_error.restype = C.c_char_p
    rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    rt.cudaGetErrorString.restype = C.c_char_p

    def ck(rc, what):
        if rc != 0:
            raise SystemExit(f"{what}: cuda error {rc} {rt.cudaGetErrorString(rc).decode()}")

    rng = np.random.default_rng(0)
    if mn:      # --mn: the ss form of the weight gradient (fbp_tc_selftest_g): out[m][n] = sum_p a[p][m] w[p][n], inputs pre-rounded to TF32 so that a single pass is exact
        tf32 = lambda v: ((v.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xffffe000)).view(np.float32)
        a = tf32(rng.standard_normal((128, 128)).astype(np.float32))
        w = tf32(rng.uniform(-1, 1, (128, 32)).astype(np.float32))
        out = np.full((128, 32), np.nan, dtype=np.float32)
    else:
        a = rng.standard_normal((128, 32)).astype(np.float32)
        w = rng.uniform(-1, 1, (32, 32)).astype(np.float32)
        out = np.full((128, 32), np.nan, dtype=np.float32)
    da, dw, do = C.c_void_p(), C.c_void_p(), C.c_void_p()
    for p, arr in ((da, a), (dw, w), (do, out)):
        ck(rt.cudaMalloc(C.byref(p), arr.nbytes), "cudaMalloc")
        ck(rt.cudaMemcpy(p, arr.ctypes.data_as(C.c_void_p), arr.nbytes, 1), "cudaMemcpy h2d")
    rc = fn(da, dw, do, variant, None)
    if rc != 0:
        raise SystemExit("fbp_tc_selftest: " + (lib.fbp_last_error() or b"?").decode())
    ck(rt.cudaDeviceSynchronize(), "cudaDeviceSynchronize")
    ck(rt.cudaMemcpy(out.ctypes.data_as(C.c_void_p), do, out.nbytes, 2), "cudaMemcpy d2h")
    if mn:
        ref = a.astype(np.float64).T @ w.astype(np.float64)
        scale = (np.abs(a).astype(np.float64).T @ np.abs(w).astype(np.float64)).max()
    else:
        ref = a.astype(np.float64) @ w.astype(np.float64).T
        scale = (np.abs(a).astype(np.float64) @ np.abs(w).astype(np.float64).T).max()
    if mn and (variant & 4):         # M = 64: find which tensor-memory lane holds each accumulator row
        lanes = [int(np.argmin(np.abs(out - ref[r]).max(axis=1))) for r in range(64)]
        errs = [float(np.abs(out[l] - ref[r]).max() / scale) for r, l in enumerate(lanes)]
        print(json.dumps({"form": "k-major ss M=64", "variant": variant, "row_to_lane": lanes, "max_err_rel": max(errs)}))
        return
    err = np.abs(out - ref)
    res = {"form": "k-major ss (weight-gradient operands)" if mn else "k-major ts", "variant": variant, "max_err_rel": float(np.nanmax(err) / scale), "nan": int(np.isnan(out).sum()),
           "bad_rows": int((err.max(axis=1) / scale > 1e-5).sum()), "bad_cols": int((err.max(axis=0) / scale > 1e-5).sum()),
           "out00": float(out[0, 0]), "ref00": float(ref[0, 0])}
    print(json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_variant(int(sys.argv[2]), mn="--mn" in sys.argv)
        sys.exit(0)
    mn = "--mn" in sys.argv
    variants = [int(v) for v in sys.argv[1:] if v != "--mn"] or ([0, 2, 4] if mn else [0, 4, 1, 2])
    lines = []
    for v in variants:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(v)] + (["--mn"] if mn else []),
                               capture_output=True, text=True, timeout=60)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else json.dumps(
                {"variant": v, "failed": (r.stderr.strip().splitlines() or ["no output"])[-1], "rc": r.returncode})
        except subprocess.TimeoutExpired:
            line = json.dumps({"variant": v, "failed": "timeout"})
        print(line, flush=True)
        lines.append(line)
    outdir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(outdir):
        with open(os.path.join(outdir, "tc_selftest.jsonl"), "a") as f:
            f.write("\n".join(lines) + "\n")
