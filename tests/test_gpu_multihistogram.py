"""data::MultiHistogram on the device (row a26, mrmd_b200_hist_*): the checks of the reference's
mrmd/data/MultiHistogram.test.cpp through the Python mirror, and gradient / smoothen / makeSymmetric / scale against the
oracle's restatement of MultiHistogram.cpp on random tables (periodic and not)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from mrmd_b200 import api as a

    assert a.L().mrmd_b200_device_count() > 0
    return a


def peak(api, a, b):
    h = api.MultiHistogram("histogram", 0.0, 10.0, 11, 2)
    d = np.zeros((11, 2))
    d[5] = (a, b)
    h.set_data(d)
    return h


def test_get_bin_and_grid(api):
    h = api.MultiHistogram("histogram", 0.0, 10.0, 10, 2)
    assert [h.getBin(v) for v in (-0.5, 0.5, 5.5, 10.5)] == [-1, 0, 5, -1]  # MultiHistogram.test.cpp:25-32
    assert h.getBinPosition(0) == 0.5 and h.getBinPosition(5) == 5.5
    assert all(h.getBin(h.getBinPosition(i)) == i for i in range(10))
    grid = api.createGrid(h)
    assert grid[0] == 0.5 and grid[5] == 5.5 and np.array_equal(grid, [h.getBinPosition(i) for i in range(10)])
    assert (h.min, h.max, h.numBins, h.numHistograms, h.binSize, h.inverseBinSize) == (0.0, 10.0, 10, 2, 1.0, 1.0)
    assert np.all(h.data == 0.0)


def test_scale_symmetric_operators(api):
    h = peak(api, 10.0, 5.0)
    h.scale(3.0)
    assert np.array_equal(h.data[5], [30.0, 15.0]) and h.data.sum() == 45.0  # :69-86
    h.scale([2.0, 4.0])
    assert np.array_equal(h.data[5], [60.0, 60.0])
    s = api.MultiHistogram("histogram", 0.0, 10.0, 10, 2)
    s.set_data(np.stack([np.arange(10.0), 10.0 - np.arange(10.0)], axis=1))
    s.makeSymmetric()
    assert np.allclose(s.data, [[4.5, 5.5]] * 10)  # :88-109
    a = peak(api, 10.0, 5.0)
    a += peak(api, 10.0, 5.0)
    assert np.array_equal(a.data[5], [20.0, 10.0])  # :135-152
    a -= peak(api, 18.0, 6.0)
    assert np.array_equal(a.data[5], [2.0, 4.0])  # :154-177
    a *= peak(api, 40.0, 2.5)
    assert np.array_equal(a.data[5], [80.0, 10.0])  # :179-202
    d = api.MultiHistogram("histogram", 0.0, 10.0, 11, 2)
    d.set_data(np.full((11, 2), 3.0))
    d /= d.copy()
    assert np.array_equal(d.data, np.ones((11, 2)))  # :204-224
    avg, cur = api.MultiHistogram("a", 0.0, 1.0, 4, 1), api.MultiHistogram("c", 0.0, 1.0, 4, 1)
    avg.set_data(np.ones((4, 1)))
    cur.set_data(np.full((4, 1), 12.0))
    api.cumulativeMovingAverage(avg, cur, 10.0)
    assert np.allclose(avg.data, 2.0)


def test_gradient_smoothen_replace(api):
    h = api.MultiHistogram("histogram", 0.0, 10.0, 10, 3)
    h.set_data(np.stack([np.ones(10), np.arange(10.0), 10.0 - np.arange(10.0)], axis=1))
    assert np.allclose(api.gradient(h, False).data, [[0.0, 1.0, -1.0]] * 10)  # :111-133
    s = api.smoothen(peak(api, 10.0, 5.0), 1.0, 3.0).data
    assert np.allclose(s[:5][::-1], s[6:])  # :226-243
    c = api.MultiHistogram("histogram", 0.0, 10.0, 11, 1)
    c.set_data(np.full((11, 1), 2.0))
    assert np.allclose(api.smoothen(c, 1.0, 3.0).data, 2.0)  # :245-263
    r = api.MultiHistogram("histogram", 0.0, 10.0, 11, 2)
    r.set_data(np.arange(11.0)[:, None] * 10 + np.arange(2.0)[None, :])
    api.replace_if_bin_position(r, api.interval_pred(-1e300, 5.0), -1.0)  # pos < 5, :265-293
    pos = np.array([r.getBinPosition(i) for i in range(11)])
    want = np.where((pos < 5.0)[:, None], -1.0, np.arange(11.0)[:, None] * 10 + np.arange(2.0)[None, :])
    assert np.array_equal(r.data, want)


@pytest.mark.parametrize("periodic", [False, True])
def test_against_oracle_on_random_tables(api, oracle, periodic):
    rng = np.random.default_rng(5)
    nb, nh = 37, 3
    table = rng.random((nb, nh)) * 4 - 1
    h = api.MultiHistogram("histogram", -2.0, 9.0, nb, nh)
    h.set_data(table)
    L = oracle.lib()
    want = np.zeros_like(table)
    L.or_hist_gradient(table.ctypes.data, want.ctypes.data, -2.0, 9.0, nb, nh, int(periodic))
    assert np.allclose(api.gradient(h, periodic).data, want, rtol=1e-14, atol=1e-14)
    L.or_hist_smoothen(table.ctypes.data, want.ctypes.data, -2.0, 9.0, nb, nh, 0.7, 2.5, int(periodic))
    assert np.allclose(api.smoothen(h, 0.7, 2.5, periodic).data, want, rtol=1e-13, atol=1e-14)
    sym = table.copy()
    L.or_hist_make_symmetric(sym.ctypes.data, nb, nh)
    h.makeSymmetric()
    assert np.array_equal(h.data, sym)
    fac = np.array([0.5, -2.0, 3.0])
    L.or_hist_scale_per_hist(sym.ctypes.data, nb, nh, fac.ctypes.data)
    h.scale(fac)
    assert np.array_equal(h.data, sym)


def test_thermodynamic_force_returns_a_multihistogram(api):
    sub = api.Subdomain([0, 0, 0], [10, 10, 10], 1.0)
    tf = api.ThermodynamicForce([1.0, 1.0], sub, 1.0, [1.0, 1.0])
    forces = np.arange(20.0).reshape(10, 2)
    tf.setForce(forces)
    h = tf.getForceHistogram()
    assert (h.min, h.max, h.numBins, h.numHistograms) == (0.0, 10.0, 10, 2) and np.array_equal(h.data, forces)
    h.scale(2.0)  # a copy: the operator's table stays
    assert np.array_equal(tf.getForce(), forces)
    assert np.all(tf.getDensityProfileHistogram().data == 0.0)
