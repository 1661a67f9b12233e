"""GPU parity: the CUDA path through the C ABI vs the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): neighbour pair sets bit-exact after canonical sorting;
per-atom forces, potential energy and virial within 1e-10 relative (fp64).
"""
import numpy as np
import pytest

from parm_b200 import workloads as W
from parity_util import backends, cpu_system, rel_err, rel_err_vec

pytestmark = pytest.mark.gpu
TOL = 1e-10

CASES = [(3, k) for k in range(4)] + [(2, k) for k in range(4)]


@pytest.mark.parametrize("ndim,kind", CASES)
def test_pairs_forces_energy(oracle_built, ndim, kind):
    from parm_b200 import sim
    w = W.random_system(700 if ndim == 3 else 500, ndim, kind, seed=11 + kind + 10 * ndim, ntypes=3, frozen=4)
    box, atoms, inter, nl, _ = sim.from_workload(w, collection=False)
    ga, gb = nl.pairs()
    atoms.reset_forces()
    gp = inter.set_forces_get_pressure(box)
    gf = atoms.peek("f")
    gE = inter.energy(box)
    gP = inter.pressure(box)
    gS = inter.stress(box)
    for be in backends(oracle_built):
        c = cpu_system(be, w, collection=False)
        ca, cb = c.pairs()
        assert len(ca) > 100
        assert np.array_equal(ga, ca) and np.array_equal(gb, cb), "pair set differs from %s" % be
        cf, cp = c.forces_and_pressure()
        assert rel_err_vec(gf, cf) < TOL
        assert rel_err(gp, cp) < TOL
        assert rel_err(gE, c.inter_energy()) < TOL
        assert rel_err(gP, c.inter_pressure()) < TOL
        assert rel_err(gS, c.inter_stress()) < TOL


@pytest.mark.parametrize("ndim,kind", [(3, 2), (3, 1), (2, 1), (3, 0), (3, 3)])
def test_verlet_trajectory(oracle_built, ndim, kind):
    """Short NVE run: same rebuild steps, same pair sets, state within tolerance."""
    from parm_b200 import sim
    w = W.random_system(600, ndim, kind, seed=5 + kind, ntypes=2, frozen=2, T=0.5)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    c = cpu_system(backends(oracle_built)[-1], w)
    collec.set_forces(True)
    c.set_forces(True)
    for block in range(4):
        collec.timestep(25)
        c.timestep(25)
        cx, cv, ca, cf = c.get_atoms()
        assert nl.which() == c.which()
        assert rel_err_vec(atoms.peek("x") - w["x"], cx - w["x"]) < 1e-9
        assert rel_err_vec(atoms.peek("v"), cv) < 1e-9
        assert rel_err_vec(atoms.peek("f"), cf) < 1e-8
        assert rel_err(collec.energy(), c.energy()) < 1e-10
        assert rel_err(collec.kinetic_energy(), c.kinetic_energy()) < 1e-10
        assert rel_err(collec.pressure(), c.pressure()) < 1e-9
        assert rel_err(collec.temp(), c.temp()) < 1e-10
    ga, gb = nl.pairs()
    ca_, cb_ = c.pairs()
    assert np.array_equal(ga, ca_) and np.array_equal(gb, cb_)


@pytest.mark.parametrize("ndim,kind", CASES)
def test_continuous_polydispersity(oracle_built, ndim, kind):
    """More distinct (eps, sigma, ...) tuples than the species table holds: parameters are gathered per
    neighbour and mixed per pair on the device (interaction.hpp:878-883, 1531-1536, 1255-1270, 970-974)."""
    from parm_b200 import sim
    w = W.random_system(600, ndim, kind, seed=70 + kind + 10 * ndim, ntypes=3, frozen=3, continuous=True, T=0.5)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    c = cpu_system(backends(oracle_built)[-1], w)
    ga, gb = nl.pairs()
    ca, cb = c.pairs()
    assert np.array_equal(ga, ca) and np.array_equal(gb, cb)
    atoms.reset_forces()
    gp = inter.set_forces_get_pressure(box)
    cf, cp = c.forces_and_pressure()
    assert rel_err_vec(atoms.peek("f"), cf) < TOL
    assert rel_err(gp, cp) < TOL
    assert rel_err(inter.energy(box), c.inter_energy()) < TOL
    assert rel_err(inter.stress(box), c.inter_stress()) < TOL
    collec.set_forces(True)
    c.set_forces(True)
    collec.timestep(40)
    c.timestep(40)
    assert nl.which() == c.which()
    assert rel_err_vec(atoms.peek("x") - w["x"], c.get_atoms()[0] - w["x"]) < 1e-9
    assert rel_err(collec.energy(), c.energy()) < 1e-10
