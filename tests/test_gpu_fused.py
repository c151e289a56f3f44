"""Fused tcgen05 MLP kernels (csrc/mlp_fused.cu): operand conventions and per-kernel parity."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(x, ref):
    return float(np.abs(x - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize("a_mode,b_mode", [(0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (2, 1)])
@pytest.mark.parametrize("a_fmt,b_fmt", [(0, 0), (1, 1)])
@pytest.mark.parametrize("N,K", [(128, 128), (32, 144), (128, 160), (16, 32), (80, 16)])
def test_umma_probe(cuda, a_mode, b_mode, a_fmt, b_fmt, N, K):
    """Slab tiles as K-major / MN-major shared-memory operands and A from tensor memory, fp16 and bf16 terms.
    (Mixed fp16 x bf16 operands in one kind::f16 instruction raise an illegal-instruction fault on sm_100a - measured.)"""
    from tensorf_b200 import _lib, ops
    rng = np.random.default_rng(1000 * a_mode + 100 * b_mode + 10 * a_fmt + b_fmt + N + K)
    A = torch.from_numpy(rng.normal(size=(128, K)).astype(np.float32)).to(cuda)
    B = torch.from_numpy((rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)).to(cuda)
    D = torch.full((128, N), float("nan"), dtype=torch.float32, device=cuda)
    _lib.check(_lib.load().tensorf_tc_umma_probe(ops._stream(), A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, a_mode, b_mode,
                                                 a_fmt, b_fmt))
    ref = A.double().cpu() @ B.double().cpu().T
    err = _rel(D.cpu().numpy(), ref.numpy())
    tol = 3e-6 if (a_fmt == 0 and b_fmt == 0) else 3e-5
    assert err < tol, f"probe a_mode={a_mode} b_mode={b_mode} fmt={a_fmt}{b_fmt} N={N} K={K}: rel err {err:.3e}"
