"""Randomness at the boundary: the hot path takes the jitter and Gumbel vectors
as INPUT arrays; keys stay in the host framework.

`render_noise(prng_key, ...)` accepts
  * a `RenderNoise` (explicit arrays — what the C ABI takes),
  * a JAX key when JAX is importable: the vectors are drawn with jax.random exactly as
    render.py:120, :158-160, :375-379 and :461-468 do,
  * a `Key` of this module: a NumPy restatement of JAX's threefry2x32 generator.  The block
    cipher is verified against the Random123 known-answer vectors (tests/test_prng.py); the
    counter layout of split/uniform/gumbel follows jax 0.9.0.1 with
    jax_threefry_partitionable=True [upstream, restated from memory — unverified against a JAX
    install because none is available in this image].
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Tuple

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds (Random123). All arguments uint32 arrays/scalars."""
    k0, k1 = np.uint32(k0), np.uint32(k1)
    x0 = np.asarray(x0, dtype=np.uint32).copy()
    x1 = np.asarray(x1, dtype=np.uint32).copy()
    ks = (k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA)))
    with np.errstate(over="ignore"):
        x0 += ks[0]
        x1 += ks[1]
        for blk in range(5):
            for r in _ROT[blk % 2]:
                x0 += x1
                x1 = (x1 << np.uint32(r)) | (x1 >> np.uint32(32 - r))
                x1 ^= x0
            x0 += ks[(blk + 1) % 3]
            x1 += ks[(blk + 2) % 3] + np.uint32(blk + 1)
    return x0, x1


@dataclasses.dataclass(frozen=True)
class Key:
    """A raw threefry key (two uint32 words), like jax.random.PRNGKey(seed)."""

    k0: int
    k1: int

    @staticmethod
    def from_seed(seed: int) -> "Key":
        return Key((seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF)


def split(key: Key, num: int = 2):
    """jax.random.split [upstream: _threefry_split_foldlike]."""
    b0, b1 = threefry2x32(key.k0, key.k1, np.zeros(num, np.uint32), np.arange(num, dtype=np.uint32))
    return [Key(int(a), int(b)) for a, b in zip(b0, b1)]


def random_bits(key: Key, shape) -> np.ndarray:
    """32 random bits per element [upstream: _threefry_random_bits_partitionable]."""
    n = int(np.prod(shape)) if len(shape) else 1
    idx = np.arange(n, dtype=np.uint64)
    b0, b1 = threefry2x32(key.k0, key.k1, (idx >> np.uint64(32)).astype(np.uint32), (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32))
    return (b0 ^ b1).reshape(shape)


def uniform(key: Key, shape, minval: float = 0.0, maxval: float = 1.0) -> np.ndarray:
    """jax.random.uniform fp32: 23 mantissa bits OR'd into 1.0, minus 1."""
    bits = random_bits(key, shape)
    floats = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo, hi = np.float32(minval), np.float32(maxval)
    return np.maximum(lo, floats * (hi - lo) + lo).astype(np.float32)


def gumbel(key: Key, shape) -> np.ndarray:
    """jax.random.gumbel: -log(-log(uniform(minval=tiny, maxval=1)))."""
    u = uniform(key, shape, minval=np.finfo(np.float32).tiny, maxval=1.0)
    return (-np.log(-np.log(u))).astype(np.float32)


@dataclasses.dataclass
class RenderNoise:
    """jitter (N,) [bounded] or (R,N) [contracted]; gumbel (N,) — shared by all rays."""

    jitter: np.ndarray
    gumbel: Optional[np.ndarray]


def render_noise(prng_key, ray_count: int, density_samples: int, contracted: bool, need_gumbel: bool = True) -> RenderNoise:
    if isinstance(prng_key, RenderNoise):
        return prng_key
    jshape: Tuple[int, ...] = (ray_count, density_samples) if contracted else (density_samples,)
    if isinstance(prng_key, Key):
        k_sample, k_rgb = split(prng_key)  # render.py:120
        return RenderNoise(uniform(k_sample, jshape), gumbel(k_rgb, (density_samples,)) if need_gumbel else None)
    try:  # a genuine JAX key
        import jax  # type: ignore

        k_sample, k_rgb = jax.random.split(prng_key)
        j = np.asarray(jax.random.uniform(k_sample, shape=jshape))
        g = np.asarray(jax.random.gumbel(k_rgb, (density_samples,))) if need_gumbel else None
        return RenderNoise(j, g)
    except ImportError as e:
        raise TypeError("prng_key must be a tensorf_b200.prng.Key, a RenderNoise, or a JAX key (JAX not importable)") from e


def render_noise_device(prng_key: Key, ray_count: int, density_samples: int, contracted: bool, device, need_gumbel: bool = True,
                        first_ray: int = 0):
    """`render_noise` drawn ON THE DEVICE by libtensorf_b200.so (`tensorf_prng_uniform/gumbel`): same keys, same
    counter layout, no host draw and no H2D copy of the (R,N) jitter.  Returns {"jitter": tensor, "gumbel": tensor|None}."""
    from . import ops
    k_sample, k_rgb = split(prng_key)  # render.py:120 (two cipher blocks: stays on the host)
    jshape: Tuple[int, ...] = (ray_count, density_samples) if contracted else (density_samples,)
    # sharded batches (first_ray = this rank's first row of the global batch): the rank draws ITS rows of the global (R,N)
    # jitter, so the ranks together reproduce the single-device draw instead of every rank reusing rows [0, R_local)
    first = first_ray * density_samples if contracted else 0
    return {"jitter": ops.prng_uniform(k_sample.k0, k_sample.k1, jshape, device, first=first),
            "gumbel": ops.prng_gumbel(k_rgb.k0, k_rgb.k1, (density_samples,), device) if need_gumbel else None}
