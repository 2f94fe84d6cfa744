"""
fbpinns_b200 — B200-native (sm_100a) implementation of the FBPINN per-step subdomain evaluation, behind the
Problem / Decomposition / Network / Scheduler / Constants / FBPINNTrainer API of benmoseley/FBPINNs.

Host language is Python; the array namespace of user plug-in code (`Problem.loss_fn`, `constraining_fn`) is
torch instead of jax.numpy because JAX is not available in this image.  All hot-path
arithmetic runs in csrc/libfbpinn_b200.so (hand-written CUDA, C ABI in include/fbpinn_b200.h); there is no
CPU fallback.
"""
__version__ = "0.1.0"
