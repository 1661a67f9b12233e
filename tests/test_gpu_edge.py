"""GPU edge cases the reference's semantics define: empty / tiny systems, atoms never add()ed,
frozen atoms, boxes too small for a cell stencil, unwrapped coordinates many images away,
the device OriginBox::diff, AtomGroup reductions, error behaviour."""
import math

import numpy as np
import pytest

from parm_b200 import workloads as W
from parity_util import backends, cpu_system, rel_err, rel_err_vec

pytestmark = pytest.mark.gpu


def test_box_diff_bit_exact_ieee_remainder():
    from parm_b200 import sim
    rng = np.random.default_rng(1)
    L = np.array([3.0, 4.5, 7.25])
    box = sim.OriginBox(L)
    atoms = sim.AtomVec(4, 1.0)
    box._attach(atoms)
    r1 = rng.uniform(-300, 300, (5000, 3))
    r2 = rng.uniform(-300, 300, (5000, 3))
    # exact ties and half-box separations
    r1[:50] = 0.0
    r2[:50] = -np.arange(50)[:, None] * (L / 2)
    got = box.diff(r1, r2)
    want = np.array([[math.remainder(a, b) for a, b in zip(p - q, L)] for p, q in zip(r1, r2)])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n", [0, 1, 2, 33])
def test_tiny_systems(oracle_built, n):
    from parm_b200 import sim
    rng = np.random.default_rng(n)
    w = dict(ndim=3, L=np.full(3, 4.0), x=rng.uniform(0, 4, (n, 3)), v=rng.standard_normal((n, 3)), m=np.ones(n),
             kind=W.KIND_LJCUT, params=np.tile([1.0, 1.0, 2.5], (n, 1)), types=np.zeros(n, np.uint32),
             eps_table=np.ones((1, 1)), skin=0.5, dt=1e-3, integrator=0)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    c = cpu_system("port", w)
    a, b = nl.pairs()
    ca, cb = c.pairs()
    assert np.array_equal(a, ca) and np.array_equal(b, cb)
    collec.set_forces(True)
    c.set_forces(True)
    collec.timestep(5)
    c.timestep(5)
    if n:
        assert rel_err_vec(atoms.peek("x"), c.get_atoms()[0]) < 1e-12
    assert collec.energy() == pytest.approx(c.energy(), rel=1e-10, abs=1e-300)


def test_small_box_below_cell_stencil(oracle_built):
    """pyparm/tests.py N=12 system: L = 3.4 < 2 (sigma + skin): no 3-cell stencil, minimum image only."""
    from parm_b200 import sim
    w = W.hertzian12()
    box, atoms, inter, nl, collec = sim.from_workload(w)
    for be in backends(oracle_built):
        c = cpu_system(be, w)
        a, b = nl.pairs()
        ca, cb = c.pairs()
        assert np.array_equal(a, ca) and np.array_equal(b, cb)
        assert rel_err(inter.energy(box), c.inter_energy()) < 1e-10


def test_partial_membership_and_frozen(oracle_built):
    from parm_b200 import sim
    w = W.random_system(400, 3, W.KIND_LJATTRACTREPULSE, seed=77, ntypes=3, frozen=25)
    rng = np.random.default_rng(3)
    w["member"] = (rng.uniform(size=400) < 0.8).astype(np.uint8)
    w["m"][rng.choice(400, 5, replace=False)] = np.inf   # isinf(m) is frozen too (collection.cpp:445)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    c = cpu_system("port", w)
    a, b = nl.pairs()
    ca, cb = c.pairs()
    assert np.array_equal(a, ca) and np.array_equal(b, cb)
    collec.set_forces(True)
    c.set_forces(True)
    collec.timestep(30)
    c.timestep(30)
    cx, cv, cacc, cf = c.get_atoms()
    assert rel_err_vec(atoms.peek("x") - w["x"], cx - w["x"]) < 1e-9
    assert rel_err_vec(atoms.peek("v"), cv) < 1e-9
    frozen = (w["m"] <= 0) | np.isinf(w["m"])
    assert np.all(atoms.peek("v")[frozen] == 0) and np.all(atoms.peek("a")[frozen] == 0)
    assert np.array_equal(atoms.peek("x")[frozen], w["x"][frozen])
    for q in ("mass", "kinetic_energy", "degrees_of_freedom"):
        assert rel_err(getattr(collec, q)() if hasattr(collec, q) else getattr(atoms, q)(), getattr(c, q)()) < 1e-12
    assert rel_err_vec(atoms.momentum(), c.momentum()) < 1e-12
    assert rel_err(collec.temp(), c.temp()) < 1e-12


def test_far_images(oracle_built):
    """Unwrapped coordinates thousands of box lengths away: the exact band widens, pairs stay exact."""
    from parm_b200 import sim
    w = W.random_system(500, 3, W.KIND_REPULSION, seed=8, ntypes=1)
    rng = np.random.default_rng(0)
    w["x"] = w["x"] + rng.integers(-5000, 5000, w["x"].shape) * w["L"]
    box, atoms, inter, nl, _ = sim.from_workload(w, collection=False)
    c = cpu_system("port", w, collection=False)
    a, b = nl.pairs()
    ca, cb = c.pairs()
    assert len(ca) > 100 and np.array_equal(a, ca) and np.array_equal(b, cb)
    atoms.reset_forces()
    inter.set_forces(box)
    assert rel_err_vec(atoms.peek("f"), c.forces()) < 1e-10


def test_velocity_helpers(oracle_built):
    from parm_b200 import sim
    w = W.lj_lattice((8, 8, 8), seed=2)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    c = cpu_system("port", w)
    collec.reset_com_velocity()
    c.reset_com_velocity()
    collec.scale_velocities_to_temp(0.7)
    c.scale_velocities_to_temp(0.7)
    assert rel_err_vec(atoms.peek("v"), c.get_atoms()[1]) < 1e-12
    collec.scale_velocities_to_energy(-2.0 * atoms.n)
    c.scale_velocities_to_energy(-2.0 * atoms.n)
    assert rel_err_vec(atoms.peek("v"), c.get_atoms()[1]) < 1e-9
    assert rel_err(collec.energy(), -2.0 * atoms.n) < 1e-9


def test_host_mirror_coherence():
    """Atom& style access between steps (LJatoms.cpp:57-60): host writes are uploaded lazily, device results
    are downloaded lazily."""
    from parm_b200 import sim
    w = W.lj_lattice((6, 6, 6), seed=4)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    collec.set_forces(True)
    collec.timestep(3)
    x3 = atoms.x.copy()
    assert not np.array_equal(x3, w["x"])
    atoms[5].v = np.array([0.1, 0.2, 0.3])           # Atom& write through the proxy
    assert np.array_equal(atoms.peek("v")[5], [0.1, 0.2, 0.3])
    collec.timestep(1)
    assert not np.array_equal(atoms.peek("x")[5], x3[5])


def test_errors():
    from parm_b200 import capi, sim
    with pytest.raises(ValueError):
        sim.CollectionSol(sim.OriginBox(5.0), sim.AtomVec(4, 1.0), 0.0, 1.0, 1.0)  # dt <= 0: invalid_argument
    box, atoms = sim.OriginBox(5.0), sim.AtomVec(4, 1.0)
    lj = sim.LJAttractRepulse(box, atoms, 0.3)
    lj.add(sim.IEpsSigCutAtom(atoms.get_id(0), [1.0, 0.5], 0, 1.0, 2.5))
    lj.add(sim.IEpsSigCutAtom(atoms.get_id(1), [0.4, 1.0], 1, 1.0, 2.5))  # asymmetric table
    with pytest.raises(ValueError):
        lj.energy(box)
    with pytest.raises(ValueError):
        lj.add(sim.IEpsSigCutAtom(atoms.get_id(1), [0.5, 1.0], 1, 1.0, 2.5))  # duplicate add
    with pytest.raises(capi.ParmUnsupported):
        sim.NeighborList(box, atoms, 0.2)  # second list on the same AtomVec


def test_drift_trigger_matches_reference_rule(oracle_built):
    """NeighborList::update_list(false) (trackers.cpp:23-53): rebuild iff the two largest displacements since the
    last rebuild sum to >= skin -- including the equality case and moves by whole box images."""
    from parm_b200 import sim
    w = W.lj_lattice((7, 7, 7), seed=9)
    w["skin"] = 0.25  # exactly representable, so 0.125 + 0.125 == skin tests the >=
    box, atoms, inter, nl, _ = sim.from_workload(w, collection=False)
    c = cpu_system("port", w, collection=False)

    def move(dx_list):
        x = w["x"].copy()
        for k, d in enumerate(dx_list):
            x[3 + 5 * k] += d
        atoms.x[:] = x
        c.set_atoms(x=x)
        return nl.update_list(False), c.update_list(False)

    base = nl.which()
    assert move([[0.1, 0, 0], [0, 0.1, 0]]) == (False, False)
    assert move([[0.125, 0, 0], [0, 0, 0.125]]) == (True, True)          # equality rebuilds
    assert nl.which() == base + 1 == c.which()
    # lastlocs were reset by the rebuild: the same positions no longer trigger
    assert (nl.update_list(False), c.update_list(False)) == (False, False)
    assert move([[0.125, 0, 0], [0, 0, 0.125], [0.2, 0, 0]]) == (False, False)   # only one atom moved since
    assert move([[0.125, 0, 0], [0, 0, 0.125], [0.2, 0, 0], [0, 0.06, 0]]) == (True, True)
    # a single atom moved by a whole box image: displacement L alone is >= skin (raw unwrapped positions)
    x = atoms.peek("x")
    x[0, 1] += w["L"][1]
    atoms.x[:] = x
    c.set_atoms(x=x)
    assert (nl.update_list(False), c.update_list(False)) == (True, True)
    a, b = nl.pairs()
    ca, cb = c.pairs()
    assert np.array_equal(a, ca) and np.array_equal(b, cb)


def test_row_capacity_growth(oracle_built):
    """Strongly non-uniform density: rows overflow the initial capacity estimate and the build retries."""
    from parm_b200 import sim
    w = W.lj_lattice((9, 9, 9), seed=12)
    x = w["x"].copy()
    x[:, 0] = x[:, 0] * 0.35  # squeeze everything into a third of the box along x
    w["x"] = x
    box, atoms, inter, nl, _ = sim.from_workload(w, collection=False)
    c = cpu_system("port", w, collection=False)
    a, b = nl.pairs()
    ca, cb = c.pairs()
    assert nl.stats()[1] > 200
    assert np.array_equal(a, ca) and np.array_equal(b, cb)
    atoms.reset_forces()
    inter.set_forces(box)
    assert rel_err_vec(atoms.peek("f"), c.forces()) < 1e-10
