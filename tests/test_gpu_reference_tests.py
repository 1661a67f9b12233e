"""The reference's own hot-path test re-expressed against the GPU path
(pyparm/tests.py:218-317, RandomHertzianVerletTest.testEnergy), plus CollectionSol statistics."""
import numpy as np
import pytest

from parm_b200 import workloads as W

pytestmark = pytest.mark.gpu


def test_random_hertzian_verlet():
    from parm_b200 import sim
    w = W.hertzian12()
    box = sim.OriginBox(w["L"])
    atoms = sim.AtomVec(w["m"])
    inter = sim.Repulsion(box, atoms, 0.4)
    for i in range(atoms.n):
        inter.add(sim.EpsSigExpAtom(atoms.get_id(i), 1.2, w["params"][i, 1], 2.0))
    for i, a in enumerate(atoms):
        a.x = w["x"][i]
        a.v = w["v"][i]
    nl = inter.neighbor_list()
    nl.update_list(True)
    collec = sim.CollectionVerlet(box, atoms, 0.01, [inter], [nl], [])
    collec.scale_velocities_to_temp(1.0)
    for _ in range(1000):
        collec.timestep()
        collec.scale_velocities_to_temp(1.0)
    lastE = collec.energy()
    EKUT = []
    for _ in range(1000):
        for _ in range(10):
            collec.timestep()
            E = collec.energy()
            assert abs(E - lastE) <= 1e-2 * abs(lastE)   # tests.py:305
            lastE = E
        EKUT.append((E, collec.kinetic_energy(), collec.potential_energy(), collec.temp()))
    E, K, U, T = np.asarray(EKUT).T
    assert abs(np.mean(T) - 1.0) < 0.1                  # tests.py:313-314
    assert abs(np.std(E) / np.mean(E)) < 1e-2           # tests.py:316-317


def test_collection_sol_thermostat_statistics():
    """Production noise (Philox + Box-Muller on the device): the Langevin thermostat must hold T and the
    Gaussian pairs must have the A&T variances/correlation (vecrand.cpp:48-85)."""
    from parm_b200 import sim
    w = W.config4(shape=(16, 16, 16))
    w.update(seed=1234, dt=0.004, damping=2.0)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    collec.set_forces(True)
    collec.timestep(2500)   # damping * t = 20: the lattice start has relaxed
    Ts = []
    for _ in range(100):
        collec.timestep(10)
        Ts.append(collec.temp())
    assert abs(np.mean(Ts) - 1.0) < 0.03
    # same seed -> same trajectory (counter based RNG), different seed -> different
    w2 = dict(w)
    b2, a2, i2, n2, c2 = sim.from_workload(w2)
    c2.set_forces(True)
    c2.timestep(20)
    w3 = dict(w, seed=99)
    b3, a3, i3, n3, c3 = sim.from_workload(w3)
    c3.set_forces(True)
    c3.timestep(20)
    b4, a4, i4, n4, c4 = sim.from_workload(w2)
    c4.set_forces(True)
    c4.timestep(20)
    assert np.array_equal(a2.peek("x"), a4.peek("x"))
    assert not np.array_equal(a2.peek("x"), a3.peek("x"))
