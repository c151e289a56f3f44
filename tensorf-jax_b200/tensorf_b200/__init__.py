"""tensorf_b200 — B200-native (sm_100a) implementation of the tensorf-jax per-ray hot path.

Host-side mirror of the reference's interface for that path (`tensor_vm`, `render`,
`networks`, `training`, `cameras`) over the C ABI in `include/tensorf_b200.h`.
The CUDA library `libtensorf_b200.so` is the product; there is no CPU fallback.
"""
from . import _lib, synthetic  # noqa: F401

__all__ = ["_lib", "ops", "synthetic", "cameras", "tensor_vm", "networks", "render", "training", "train_config"]
