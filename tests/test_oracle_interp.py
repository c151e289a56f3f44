"""Pin the oracle's map_coordinates restatement (tensor_vm.py:226-250) against SciPy's
`scipy.ndimage.map_coordinates(order=1, mode="nearest")`, the routine that
jax.scipy.ndimage.map_coordinates re-implements, plus hand-checkable micro cases."""
import numpy as np
import scipy.ndimage
import torch

import tensorf_oracle as O


def test_map_coordinates_1d_2d_vs_scipy():
    rng = np.random.default_rng(0)
    for G in (2, 5, 16):
        grid1 = rng.normal(size=(3, G))
        grid2 = rng.normal(size=(3, G, G))
        x = rng.uniform(-1.5, G + 0.5, size=(2, 400))
        x[:, :6] = np.array([[0.0, G - 1.0, -0.25, G - 0.5, 1.0, G - 1.0], [G - 1.0, 0.0, 0.5, -3.0, 1.0, G - 1.0]])
        out1 = O.linear_interpolation_with_channel_axis(torch.from_numpy(grid1), torch.from_numpy(x[:1])).numpy()
        out2 = O.linear_interpolation_with_channel_axis(torch.from_numpy(grid2), torch.from_numpy(x)).numpy()
        for c in range(3):
            ref1 = scipy.ndimage.map_coordinates(grid1[c], x[:1], order=1, mode="nearest")
            ref2 = scipy.ndimage.map_coordinates(grid2[c], x, order=1, mode="nearest")
            np.testing.assert_allclose(out1[c], ref1, rtol=0, atol=1e-12)
            np.testing.assert_allclose(out2[c], ref2, rtol=0, atol=1e-12)


def test_fp32_instance_close_to_scipy():
    rng = np.random.default_rng(1)
    G = 128
    grid = rng.normal(size=(2, G, G)).astype(np.float32)
    x = rng.uniform(0, G - 1, size=(2, 5000)).astype(np.float32)
    out = O.linear_interpolation_with_channel_axis(torch.from_numpy(grid), torch.from_numpy(x)).numpy()
    ref = scipy.ndimage.map_coordinates(grid[0].astype(np.float64), x.astype(np.float64), order=1, mode="nearest")
    assert np.abs(out[0] - ref).max() < 5e-6


def test_vm_single_g2_hand_case():
    # G=2, C=1: vector [1, 3], matrix [[1, 2], [3, 4]]; at ijk=(0,0,0) -> x=0.5 everywhere:
    # lin = 2, bil = 2.5, product 5. At ijk=(-1, 1, -1): lin=v[0]=1, bil=m[1,0]=3.
    vec = torch.tensor([[1.0, 3.0]], dtype=torch.float64)
    mat = torch.tensor([[[1.0, 2.0], [3.0, 4.0]]], dtype=torch.float64)
    ijk = torch.tensor([[0.0, -1.0], [0.0, 1.0], [0.0, -1.0]], dtype=torch.float64)
    out = O.vm_single_interpolate(vec, mat, ijk)
    assert out.tolist() == [[5.0, 3.0]]


def test_vm_axis_permutations():
    # tensor_vm.py:50-52: pair 0 (line x; plane y,z), pair 1 (line z; plane x,y), pair 2 (line y; plane z,x)
    G = 3
    vec = torch.zeros(3, 1, G, dtype=torch.float64)
    mat = torch.zeros(3, 1, G, G, dtype=torch.float64)
    vec[:, 0] = torch.tensor([1.0, 10.0, 100.0])          # value identifies the line index
    for a in range(G):
        for b in range(G):
            mat[:, 0, a, b] = 1000.0 * (a + 1) + 10000.0 * (b + 1) * 0 + (b + 1)  # row a, col b
    # point at integer grid coords (x,y,z) = (0,1,2)  -> ijk = coord/(G-1)*2-1
    ijk = torch.tensor([[0.0], [1.0], [2.0]], dtype=torch.float64) / (G - 1) * 2 - 1
    out = O.vm_interpolate(vec, mat, ijk).reshape(3)
    # pair 0: line[x=0]=1, plane[row=y=1, col=z=2] = 2000+3
    # pair 1: line[z=2]=100, plane[row=x=0, col=y=1] = 1000+2
    # pair 2: line[y=1]=10, plane[row=z=2, col=x=0] = 3000+1
    assert out.tolist() == [1 * 2003.0, 100 * 1002.0, 10 * 3001.0]


def test_out_of_range_collapses_to_edge():
    vec = torch.tensor([[2.0, 4.0, 8.0]], dtype=torch.float64)
    out = O.linear_interpolation_with_channel_axis(vec, torch.tensor([[-5.0, -0.5, 2.5, 7.0]], dtype=torch.float64))
    assert out.tolist() == [[2.0, 2.0, 8.0, 8.0]]
