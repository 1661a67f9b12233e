"""GPU at benchmark sizes: parity with the oracle where the oracle finishes in seconds (64k atoms),
size-independent properties at BASELINE.json's full single-GPU size (1e6 atoms)."""
import numpy as np
import pytest

from parm_b200 import workloads as W
from parity_util import cpu_system, rel_err, rel_err_vec

pytestmark = pytest.mark.gpu


def test_config3_shape_64k_vs_oracle(oracle_built):
    from parm_b200 import sim
    w = W.lj_lattice((40, 40, 40), seed=3003)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    c = cpu_system("port", w, injected=True)
    a, b = nl.pairs()
    ca, cb = c.pairs()
    assert np.array_equal(a, ca) and np.array_equal(b, cb)
    collec.set_forces(True)
    c.set_forces(True)
    assert rel_err_vec(atoms.peek("f"), c.get_atoms()[3]) < 1e-10
    assert rel_err(collec.potential_energy(), c.potential_energy()) < 1e-10
    assert rel_err(collec.virial(), c.virial()) < 1e-10
    collec.timestep(12)
    c.timestep(12)
    assert nl.which() == c.which()
    assert rel_err_vec(atoms.peek("x") - w["x"], c.get_atoms()[0] - w["x"]) < 1e-10
    a, b = nl.pairs()
    ca, cb = c.pairs()
    assert np.array_equal(a, ca) and np.array_equal(b, cb)


def test_config2_2d_harmonic_100k_vs_oracle(oracle_built):
    from parm_b200 import sim
    w = W.config2()
    box, atoms, inter, nl, collec = sim.from_workload(w)
    c = cpu_system("port", w, injected=True)
    a, b = nl.pairs()
    ca, cb = c.pairs()
    assert np.array_equal(a, ca) and np.array_equal(b, cb)
    collec.set_forces(True)
    c.set_forces(True)
    assert rel_err_vec(atoms.peek("f"), c.get_atoms()[3]) < 1e-10
    assert rel_err(collec.potential_energy(), c.potential_energy()) < 1e-10
    collec.timestep(20)
    c.timestep(20)
    assert rel_err(collec.energy(), c.energy()) < 1e-10


def test_config3_full_size_properties():
    """N = 1e6 (BASELINE configs[2]): list symmetric (every pair listed from both ends), Newton's third law
    (sum f = 0), NVE energy drift small, momentum conserved."""
    from parm_b200 import sim
    w = W.config3()
    box, atoms, inter, nl, collec = sim.from_workload(w)
    n = atoms.n
    mean, mx = nl.stats()
    assert 95 < mean < 125
    a, b = nl.pairs()  # download_pairs fails if the full list is not symmetric
    assert len(a) == nl.numpairs() and np.all(a > b)
    collec.set_forces(True)
    fsum = atoms.com_force()
    fscale = np.sqrt((atoms.peek("f") ** 2).sum(1)).max()
    assert np.abs(fsum).max() < 1e-9 * fscale * np.sqrt(n)
    p0 = atoms.momentum()
    E0 = collec.energy()
    collec.timestep(200)
    E1 = collec.energy()
    assert abs(E1 - E0) < 1e-3 * abs(E0)   # dt = 0.004 with a force discontinuity at the cut: see the drift test
    assert np.abs(atoms.momentum() - p0).max() < 1e-7 * n ** 0.5
    assert collec.stats()["rebuilds"] >= 10


def test_nve_drift_matches_reference_10k_steps(oracle_built):
    """BASELINE config 1 (LJatoms.cpp-like, N=1000): 10^4 NVE steps on the GPU and on the CPU oracle from the
    same inputs. Trajectories decorrelate after ~10^3 steps (chaos), so the comparison is statistical:
    energy-drift envelope and neighbour-list rebuild count."""
    from parm_b200 import sim
    from parity_util import backends
    w = W.config1()
    box, atoms, inter, nl, collec = sim.from_workload(w)
    c = cpu_system(backends(oracle_built)[-1], w)
    for s_ in (collec, c):
        s_.reset_com_velocity()
        s_.scale_velocities_to_energy(w["x"].shape[0] / 4.0)   # LJatoms.cpp:85-86
        s_.set_forces(True)
    E0g, E0c = collec.energy(), c.energy()
    assert rel_err(E0g, E0c) < 1e-10
    dg, dc = [], []
    for block in range(100):
        collec.timestep(100)
        c.timestep(100)
        dg.append(abs(collec.energy() - E0g) / abs(E0g))
        dc.append(abs(c.energy() - E0c) / abs(E0c))
        if block == 2:   # still correlated after 300 steps
            assert rel_err_vec(atoms.peek("x") - w["x"], c.get_atoms()[0] - w["x"]) < 1e-7
    dg, dc = np.array(dg), np.array(dc)
    assert dg.max() < 1e-5 and dc.max() < 1e-5
    assert dg.max() < 3 * dc.max() + 1e-9 and dc.max() < 3 * dg.max() + 1e-9
    assert abs(nl.which() - c.which()) <= max(2, 0.05 * c.which())


def test_peak_probes_are_plausible():
    """csrc/probe.cu: the DFMA-chain and streaming-copy probes bench.py uses as roofline denominators.
    B200: fp64 vector peak ~37 TFLOP/s, HBM3e copy ~6.5 TB/s; anything far outside means a broken probe."""
    import ctypes as C
    from parm_b200 import capi
    f, c = C.c_double(), C.c_double()
    capi.call("parm_b200_probe_peaks", 0, C.byref(f), C.byref(c))
    assert 1.5e4 < f.value < 6e4, f.value
    assert 3.0e3 < c.value < 9.0e3, c.value
