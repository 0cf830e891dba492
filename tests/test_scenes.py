"""Host-side scene helpers: splitmix64 check values and the synthetic cloth recipe."""
import numpy as np

from collisiondetection_b200 import scenes


def test_splitmix64_check_values():
    got = scenes.splitmix64(np.uint64(20261017) + np.arange(3, dtype=np.uint64))
    assert [int(x) for x in got] == [0x7066B371864289D7, 0xB071EAD408738983, 0x52A8EFCFB33C2D7B]


def test_cloth_shapes_and_faces():
    n = 21
    q0, q1, f, eta = scenes.cloth(n)
    assert q0.shape == (n * n, 3) and q1.shape == q0.shape
    assert f.shape == (2 * (n - 1) ** 2, 3) and f.min() == 0 and f.max() == n * n - 1
    assert eta == 0.01 / (n - 1)
    # x,y unchanged by the motion except for the +-0.1h noise
    h = 1.0 / (n - 1)
    assert np.abs(q1[:, :2] - q0[:, :2]).max() <= 0.1 * h * (1 + 1e-12)
    # first triangle of cell (0,0) ((i+j) even): (i0,i1,i2); of cell (0,1) (odd): (i0,i1,i3)
    assert f[0].tolist() == [0, n, n + 1]
    assert f[1].tolist() == [1, n + 1, 2]


def test_cloth_full_size_counts():
    # only the index arithmetic: n=1415 -> 3,998,792 triangles (BASELINE config C5)
    n = 1415
    assert n * n == 2002225 and 2 * (n - 1) ** 2 == 3998792


def test_cloth_strip_is_a_sample_of_the_same_cloth():
    n = 41
    q0, q1, f, eta = scenes.cloth(n)
    s0, s1, sf, seta = scenes.cloth(n, cols=(10, 20))
    assert seta == eta
    keep = (np.arange(n * n) % n >= 10) & (np.arange(n * n) % n < 20)
    assert np.array_equal(s0, q0[keep]) and np.array_equal(s1, q1[keep])
    assert sf.shape == (2 * 9 * (n - 1), 3) and sf.max() == 10 * n - 1
