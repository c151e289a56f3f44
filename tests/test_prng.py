"""Host PRNG restatement: threefry2x32 against the Random123 known-answer vectors
(SURVEY.md §8f rank 3), and structural properties of split/uniform/gumbel."""
import numpy as np

from tensorf_b200 import prng, render, synthetic as S
import tensorf_oracle as O

KAT = [
    ((0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6B200159, 0x99BA4EFE)),
    ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
    ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0)),
]


def test_threefry_known_answers():
    for key, ctr, out in KAT:
        a, b = prng.threefry2x32(key[0], key[1], np.uint32(ctr[0]), np.uint32(ctr[1]))
        assert (int(a), int(b)) == out


def test_split_uniform_gumbel_properties():
    k = prng.Key.from_seed(0)
    a, b = prng.split(k)
    # jax.random.split(jax.random.key(0)) with threefry_partitionable (default since jax 0.5)
    assert (a.k0, a.k1) == (1797259609, 2579123966) and (b.k0, b.k1) == (928981903, 3453687069)
    u = prng.uniform(a, (1000, 7))
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02
    assert np.array_equal(prng.uniform(a, (7000,)), u.reshape(-1))      # counter = flat row-major index
    g = prng.gumbel(b, (5000,))
    assert np.isfinite(g).all() and abs(g.mean() - 0.5772) < 0.05
    n1 = prng.render_noise(k, 16, 32, contracted=False)
    n2 = prng.render_noise(k, 16, 32, contracted=True)
    assert n1.jitter.shape == (32,) and n2.jitter.shape == (16, 32) and n1.gumbel.shape == (32,)
    assert np.array_equal(n1.gumbel, n2.gumbel)                          # same rgb key (render.py:120)
    assert prng.render_noise(n1, 1, 1, False) is n1


def test_contracted_schedule_matches_oracle():
    for n in (8, 41, 665, 1558):
        b1, d1 = render.contracted_schedule(0.05, 200.0, n)
        b2, d2 = O.contracted_schedule(0.05, 200.0, n)
        assert np.array_equal(b1, b2) and np.array_equal(d1, d2)
        assert b1.dtype == np.float32 and (d1 > 0).all()
    b, d = render.contracted_schedule(0.05, 200.0, 665)
    assert abs(d[:300].mean() - 0.0030) < 2e-4 and 11.0 < d[-1] < 13.5


def test_workload_byte_model():
    assert S.lego_workload().fwd_bytes_per_ray() == 368640 and S.lego_workload().train_bytes_per_ray() == 1105920
    assert S.lego_workload(N=256, K=38).train_bytes_per_ray() == 1278720
    assert S.lego_workload(R=16384, G=300).train_bytes_per_ray() == 2592000
    assert S.dozer_workload().train_bytes_per_ray() == 5622912
    assert S.render360_workload().fwd_bytes_per_ray() == 1032192
