"""CPU-only: the synthetic input generators."""
import numpy as np

from parm_b200 import workloads as W


def _temp(w):
    m, v = w["m"], w["v"]
    n, d = v.shape
    vc = (v * m[:, None]).sum(0) / m.sum()
    return (m[:, None] * (v - vc) ** 2).sum() / (d * n - d)


def test_lj_lattice_state_point():
    w = W.lj_lattice((10, 10, 10))
    assert w["x"].shape == (1000, 3)
    assert abs(1000 / np.prod(w["L"]) - 1.1939) < 1e-12
    assert abs(_temp(w) - 1.44) < 1e-12
    assert np.abs((w["v"] * w["m"][:, None]).sum(0)).max() < 1e-9


def test_config2_is_2d_bidisperse():
    w = W.config2(nx=25, ny=40)
    assert w["ndim"] == 2 and w["x"].shape == (1000, 2)
    s = w["params"][:, 1]
    assert set(np.unique(s)) == {1.0, 1.4} and abs((s == 1.0).mean() - 0.5) < 1e-12
    phi = (np.pi * s ** 2 / 4).sum() / np.prod(w["L"])
    assert abs(phi - 0.9) < 1e-12


def test_config4_density():
    w = W.config4(shape=(6, 6, 6))
    s = w["params"][:, 1]
    phi = (np.pi / 6 * s ** 3).sum() / np.prod(w["L"])
    assert abs(phi - 0.55) < 1e-12 and w["integrator"] == W.SOL


def test_hertzian12_matches_reference_setup():
    w = W.hertzian12()
    assert w["x"].shape == (12, 3) and np.all(w["x"] >= 0) and np.all(w["x"] < w["L"][0])
    assert np.allclose(w["m"], w["params"][:, 1] ** 3)


def test_random_system_is_deterministic_and_discrete():
    a, b = W.random_system(300, 3, 2, seed=5), W.random_system(300, 3, 2, seed=5)
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["params"], b["params"])
    keys = {tuple(r) for r in np.column_stack([a["params"], a["types"]])}
    assert len(keys) <= 32
