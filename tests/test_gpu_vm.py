"""CUDA TensorVM.interpolate (tensor_vm.py:42-89) vs the CPU oracle, through the C ABI."""
import numpy as np
import pytest
import torch

import tensorf_oracle as O
from helpers import T, assert_close_grad, assert_close_out

pytestmark = pytest.mark.gpu


def _factors(C, G, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.normal(0, 0.1, (3, C, G)).astype(np.float32), rng.normal(0, 0.1, (3, C, G, G)).astype(np.float32))


@pytest.mark.parametrize("C,G", [(1, 2), (3, 5), (16, 33), (48, 17), (5, 128)])
def test_pack_unpack_roundtrip(cuda, C, G):
    from tensorf_b200 import ops
    v, m = _factors(C, G)
    packed = ops.vm_pack(T(v, device=cuda), T(m, device=cuda))
    v2, m2 = ops.vm_unpack(packed, C, G)
    assert np.array_equal(v2.cpu().numpy(), v) and np.array_equal(m2.cpu().numpy(), m)
    # packed layout: planes[P][a][b][c] after lines[P][i][c], channels padded to a multiple of 4
    Cp = (C + 3) // 4 * 4
    pk = packed.cpu().numpy()
    lines = pk[: 3 * G * Cp].reshape(3, G, Cp)
    planes = pk[3 * G * Cp:].reshape(3, G, G, Cp)
    assert np.array_equal(lines[..., :C], v.transpose(0, 2, 1))
    assert np.array_equal(planes[..., :C], m.transpose(0, 2, 3, 1))
    assert not lines[..., C:].any() and not planes[..., C:].any()


@pytest.mark.parametrize("C,G,B", [(1, 2, 7), (2, 9, 1000), (16, 128, 5000), (48, 31, 777), (6, 16, 33)])
@pytest.mark.parametrize("feature_major", [False, True])
def test_interp_fwd(cuda, C, G, B, feature_major):
    from tensorf_b200 import ops
    v, m = _factors(C, G, seed=C + G)
    rng = np.random.default_rng(B)
    ijk = rng.uniform(-1.0, 1.0, (3, B)).astype(np.float32)
    # edge cases: out of range, exact endpoints, exact integer grid coordinates
    ijk[:, 0] = [-1.0, 1.0, 0.0]
    if B > 4:
        ijk[:, 1] = [1.5, -1.25, 1.0]
        ijk[:, 2] = [-1.0, -1.0, -1.0]
        ijk[:, 3] = [1.0, 1.0, 1.0]
        ijk[:, 4] = (np.array([0, G - 1, G // 2], dtype=np.float32) / (G - 1)) * 2 - 1
    ref = O.vm_interpolate(T(v), T(m), T(ijk)).numpy()
    packed = ops.vm_pack(T(v, device=cuda), T(m, device=cuda))
    out = ops.vm_interp_fwd(packed, T(ijk, device=cuda), C, G, feature_major).cpu().numpy()
    if feature_major:
        out = out.T
    assert_close_out(out, ref, rtol=1e-5, what="vm_interp_fwd")


@pytest.mark.parametrize("C,G,B", [(2, 9, 500), (16, 64, 4000), (7, 12, 300)])
def test_interp_bwd(cuda, C, G, B):
    from tensorf_b200 import ops
    v, m = _factors(C, G, seed=3)
    rng = np.random.default_rng(5)
    ijk = rng.uniform(-1.05, 1.05, (3, B)).astype(np.float32)
    dout = rng.normal(size=(3 * C, B)).astype(np.float32)
    vt, mt = T(v, torch.float64).requires_grad_(True), T(m, torch.float64).requires_grad_(True)
    (O.vm_interpolate(vt, mt, T(ijk, torch.float64)) * T(dout, torch.float64)).sum().backward()
    packed = ops.vm_pack(T(v, device=cuda), T(m, device=cuda))
    dpk = ops.vm_interp_bwd(packed, T(ijk, device=cuda), T(dout, device=cuda), C, G, False)
    dv, dm = ops.vm_unpack(dpk, C, G)
    assert_close_grad(dv.cpu().numpy(), vt.grad.numpy(), what="d vector")
    assert_close_grad(dm.cpu().numpy(), mt.grad.numpy(), what="d matrix")


def test_interp_rejects_bad_args(cuda):
    from tensorf_b200 import ops
    from tensorf_b200._lib import TensorfError
    v, m = _factors(2, 4)
    packed = ops.vm_pack(T(v, device=cuda), T(m, device=cuda))
    with pytest.raises(ValueError):
        ops.vm_interp_fwd(packed, torch.zeros(2, 5, device=cuda), 2, 4)
    with pytest.raises(ValueError):
        ops.vm_interp_fwd(packed, torch.zeros(3, 5), 2, 4)  # CPU tensor: no CPU path
    with pytest.raises(TensorfError):
        ops.topk_select(torch.zeros(4, 8, device=cuda), 9)  # K > N
