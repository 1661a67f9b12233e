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


def test_config3_full_size_1m_vs_oracle(oracle_built):
    """BASELINE configs[2] at its FULL size (N = 1e6) against the reference's own NListed / CollectionVerlet code
    (oracle/_ref when built, else the C port) fed by the harness cell list (InjectedNeighborList, SURVEY 8c
    "Large-N oracle"): pair set bit-exact, forces / energy / virial within 1e-10, then 8 steps across a rebuild
    (interaction.hpp:2102-2115, 2166-2175; collection.cpp:442-469; trackers.cpp:19-85). ~1-2 min of one host core."""
    from parm_b200 import sim
    from parity_util import backends
    w = W.config3()
    be = backends(oracle_built)[-1]
    box, atoms, inter, nl, collec = sim.from_workload(w)
    assert nl.tile_stats()[0]  # the cell-tile pair kernel is the one under test
    c = cpu_system(be, w, injected=True)
    a, b = nl.pairs()
    ca, cb = c.pairs()
    assert len(a) == len(ca) > 45_000_000
    assert np.array_equal(a, ca) and np.array_equal(b, cb)
    del a, b, ca, cb
    collec.set_forces(True)
    c.set_forces(True)
    assert rel_err_vec(atoms.peek("f"), c.get_atoms()[3]) < 1e-10
    assert rel_err(collec.potential_energy(), c.potential_energy()) < 1e-10
    assert rel_err(collec.virial(), c.virial()) < 1e-10
    assert rel_err(collec.pressure(), c.pressure()) < 1e-10
    collec.timestep(8)
    c.timestep(8)
    assert c.which() >= 2 and nl.which() == c.which()  # at least one drift-triggered rebuild, at the same step
    x, v, _, f = c.get_atoms()
    assert rel_err_vec(atoms.peek("x") - w["x"], x - w["x"]) < 1e-10
    assert rel_err_vec(atoms.peek("v"), v) < 1e-10
    assert rel_err_vec(atoms.peek("f"), f) < 1e-10
    assert rel_err(collec.energy(), c.energy()) < 1e-10
    a, b = nl.pairs()
    ca, cb = c.pairs()
    assert np.array_equal(a, ca) and np.array_equal(b, cb)


def test_config4_shape_256k_sol_vs_oracle(oracle_built):
    """BASELINE configs[3] (binary WCA-like LJRepulsePair under CollectionSol, collection.cpp:265-322) at 64^3 =
    262 144 atoms with INJECTED normals (the Boost stream is unpinned): pair set bit-exact, forces 1e-10,
    trajectories after 20 Langevin steps across a drift-triggered rebuild."""
    from parm_b200 import sim
    from parity_util import backends
    w = W.config4(shape=(64, 64, 64), seed=4004)
    be = backends(oracle_built)[-1]
    steps = 20
    box, atoms, inter, nl, collec = sim.from_workload(w)
    c = cpu_system(be, w, injected=True)
    a, b = nl.pairs()
    ca, cb = c.pairs()
    assert np.array_equal(a, ca) and np.array_equal(b, cb)
    collec.set_forces(True)
    c.set_forces(True)
    assert rel_err_vec(atoms.peek("f"), c.get_atoms()[3]) < 1e-10
    assert rel_err(collec.potential_energy(), c.potential_energy()) < 1e-10
    z = np.random.default_rng(44).standard_normal((steps, w["x"].shape[0], 2, 3))
    collec.inject_noise(z)
    c.inject_noise(z)
    collec.timestep(steps)
    c.timestep(steps)
    x, v, _, f = c.get_atoms()
    assert c.which() >= 2 and nl.which() == c.which()
    assert rel_err_vec(atoms.peek("x") - w["x"], x - w["x"]) < 1e-10
    assert rel_err_vec(atoms.peek("v"), v) < 1e-10
    assert rel_err_vec(atoms.peek("f"), f) < 1e-9
    a, b = nl.pairs()
    ca, cb = c.pairs()
    assert np.array_equal(a, ca) and np.array_equal(b, cb)


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
