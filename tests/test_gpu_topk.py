"""Selection stage (render.py:461-469): CUDA radix select vs the oracle, bit-exact index sets."""
import ctypes
import pathlib

import numpy as np
import pytest
import torch

import tensorf_oracle as O
from helpers import T

pytestmark = pytest.mark.gpu


def _check(cuda, g, K):
    from tensorf_b200 import ops
    idx = ops.topk_select(T(g, device=cuda), K).cpu().numpy()
    ref = np.sort(O.gumbel_topk(T(g), K).numpy(), axis=-1)
    assert idx.dtype == np.int32 and idx.shape == ref.shape
    assert (np.diff(idx, axis=-1) > 0).all() if K > 1 else True  # ascending, unique
    assert np.array_equal(idx, ref)


@pytest.mark.parametrize("R,N,K", [(1, 1, 1), (5, 7, 7), (64, 37, 5), (300, 221, 33), (17, 519, 77), (9, 665, 99),
                                   (6, 1558, 233), (33, 512, 128), (4, 2048, 1)])
def test_random(cuda, R, N, K):
    rng = np.random.default_rng(N * 1000 + K)
    _check(cuda, rng.normal(size=(R, N)).astype(np.float32) * 5, K)


def test_ties_lower_index_first(cuda):
    g = np.zeros((4, 100), np.float32)            # all equal: take the first K
    _check(cuda, g, 10)
    g = np.tile(np.array([1.0, 2.0, 2.0, 1.0, 2.0, 0.0, 2.0], np.float32), (3, 1))
    for K in range(1, 8):
        _check(cuda, g, K)
    rng = np.random.default_rng(0)
    g = rng.integers(-3, 3, size=(50, 221)).astype(np.float32)  # heavy ties
    _check(cuda, g, 33)


def test_neg_inf_runs_and_signed_zero(cuda):
    rng = np.random.default_rng(1)
    g = rng.normal(size=(20, 221)).astype(np.float32)
    g[:, 40:] = -np.inf                       # log(p_terminates = 0)
    _check(cuda, g, 33)                       # fewer finite than K? 40 >= 33 ok
    _check(cuda, g, 60)                       # must take 20 of the -inf run, lowest indices
    g2 = np.full((2, 64), -np.inf, np.float32)
    _check(cuda, g2, 5)
    g3 = np.array([[0.0, -0.0, 0.0, -0.0, 1e-45, -1e-45]], np.float32)
    # XLA's TopK key orders -0 < +0; the stable float sort treats them as equal. Only check the
    # unambiguous prefix.
    from tensorf_b200 import ops
    idx = ops.topk_select(T(g3, device=cuda), 1).cpu().numpy()
    assert idx.tolist() == [[4]]


def test_matches_c_oracle(cuda):
    """The plain-C restatement (oracle/topk_select.c) agrees too (independent of torch.sort)."""
    so = pathlib.Path(__file__).resolve().parents[1] / "oracle" / "_build" / "liboracle_topk.so"
    if not so.exists():
        pytest.skip("oracle/_build/liboracle_topk.so not built")
    lib = ctypes.CDLL(str(so))
    rng = np.random.default_rng(3)
    g = np.round(rng.normal(size=(40, 300)) * 4).astype(np.float32)
    K = 45
    ref = np.empty((40, K), np.int32)
    lib.oracle_topk_select(g.ctypes.data_as(ctypes.c_void_p), 40, 300, K, ref.ctypes.data_as(ctypes.c_void_p))
    from tensorf_b200 import ops
    idx = ops.topk_select(T(g, device=cuda), K).cpu().numpy()
    assert np.array_equal(idx, np.sort(ref, axis=-1))
