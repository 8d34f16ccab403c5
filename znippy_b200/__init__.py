"""znippy_b200 — B200 (sm_100a) back-end for znippy's per-chunk codec + integrity hot path.

    znippy_b200.codec     mirror of znippy-common/src/codec.rs + blake3::hash, batch-first
    znippy_b200.archive   `.znippy` v0.7 container and the read / random-access / write loop shells
    znippy_b200._native   ctypes binding of libznippy_cuda.so (include/znippy_cuda.h)

Everything that computes runs CUDA kernels from libznippy_cuda.so; importing this package never falls back to a
CPU implementation."""
from . import _native  # noqa: F401
from ._native import Ctx, NativeError, Plan, default_ctx  # noqa: F401

__all__ = ["Ctx", "Plan", "NativeError", "default_ctx"]
