"""CollectionDamped, SolHT, Overdamped, NoseHoover, GaussianT and Gear3A-6A on the GPU against
the oracle (the compiled reference when present, else its C restatement): trajectories, rebuild counts, energies
and thermostat state, through multi-step calls (steps queued with the speculative rebuild guard) and one step
at a time."""
import numpy as np
import pytest

from parm_b200 import workloads as W
from parity_util import backends, cpu_system, rel_err, rel_err_vec

pytestmark = pytest.mark.gpu
CASES = sorted(W.INTEGRATOR_CASES)
IDS = [W.INTEGRATOR_CASES[k][0] for k in CASES]


def lattice_case(integ, ndim):
    """A few thousand atoms on a jittered lattice: enough steps for several neighbour-list rebuilds."""
    if ndim == 3:
        w = W.lj_lattice((12, 12, 12), seed=60 + integ, T=1.0)
    else:
        w = W.config2(nx=60, ny=60, seed=60 + integ)
        w["v"] = w["v"] * 20.0
    params = W.INTEGRATOR_CASES[integ][1]
    if integ == 4 and ndim == 3:
        params = (0.05,)  # explicit Euler on the LJ lattice: gamma*dt*stiffness must stay below 2 or it blows up
    w.update(integrator=integ, integ_params=params)
    return w


def noise_for(w, steps, integ):
    if integ != 3:
        return None
    return np.random.default_rng(9).standard_normal((steps, int((w["m"] > 0).sum()), w["ndim"]))


def compare(collec, atoms, nl, s, w, tol_x=1e-8):
    x, v, a, f = s.get_atoms()
    assert nl.which() == s.which()
    assert rel_err_vec(atoms.peek("x") - w["x"], x - w["x"]) < tol_x
    assert rel_err_vec(atoms.peek("v"), v) < tol_x
    assert rel_err_vec(atoms.peek("a"), a) < 1e-7
    assert rel_err(collec.energy(), s.energy()) < 1e-9
    assert rel_err(collec.pressure(), s.pressure()) < 1e-8


@pytest.mark.parametrize("integ", CASES, ids=IDS)
@pytest.mark.parametrize("ndim", [3, 2])
def test_integrator_matches_oracle(oracle_built, integ, ndim):
    from parm_b200 import sim
    w = lattice_case(integ, ndim)
    steps = 120
    be = backends(oracle_built)[-1]
    box, atoms, inter, nl, collec = sim.from_workload(w)
    s = cpu_system(be, w, injected=True)
    collec.set_forces(True)
    s.set_forces(True)
    z = noise_for(w, steps, integ)
    if z is not None:
        collec.inject_noise(z)
        s.inject_noise(z)
    collec.timestep(steps)
    s.timestep(steps)
    compare(collec, atoms, nl, s, w)
    if integ != 4:  # (overdamped relaxation barely moves the atoms)
        assert nl.which() > 1  # at least one rebuild went through the guarded pipeline
    if integ == 5:
        assert rel_err(collec._scalars(), s.get_scalars()) < 1e-9
        assert abs(collec.get_xi()) > 0


@pytest.mark.parametrize("integ", CASES, ids=IDS)
def test_integrator_ragged_system_step_by_step(oracle_built, integ):
    """Random ragged systems (vacancies, unwrapped coordinates, several species, frozen atoms where the integrator
    tests for them), advanced one timestep() call at a time."""
    from parm_b200 import sim
    w = W.integrator_system(integ, ndim=3, n=1500, seed=90 + integ)
    steps = 25
    be = backends(oracle_built)[-1]
    box, atoms, inter, nl, collec = sim.from_workload(w)
    s = cpu_system(be, w, injected=True)
    collec.set_forces(True)
    s.set_forces(True)
    z = noise_for(w, steps, integ)
    if z is not None:
        collec.inject_noise(z)
        s.inject_noise(z)
    for _ in range(steps):
        collec.timestep()
    s.timestep(steps)
    compare(collec, atoms, nl, s, w, tol_x=1e-7)


def test_solht_thermostat_reaches_temperature():
    """Production noise (Philox, not injected): the Honeycutt-Thirumalai Langevin integrator equilibrates to T."""
    from parm_b200 import sim
    w = W.config4(shape=(14, 14, 14), seed=5)
    w.update(integrator=3, integ_params=(2.0, 1.0), dt=0.002)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    collec.set_forces(True)
    collec.timestep(4000)
    Ts = []
    for _ in range(20):
        collec.timestep(100)
        Ts.append(collec.temp())
    assert abs(np.mean(Ts) - 1.0) < 0.05


def test_nosehoover_hamiltonian_is_conserved():
    from parm_b200 import sim
    w = W.lj_lattice((10, 10, 10), seed=11, T=1.0)
    w.update(integrator=5, integ_params=(50.0, 1.0), dt=0.002)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    collec.set_forces(True)
    H0 = collec.hamiltonian()
    collec.timestep(2000)
    H1 = collec.hamiltonian()
    assert abs(H1 - H0) / abs(H0) < 2e-3
    assert collec.get_lns() != 0.0


def test_extra_integrators_refuse_sharded_contexts_and_bad_types():
    import ctypes as C
    from parm_b200 import capi, sim
    atoms = sim.AtomVec(np.ones(8), ndim=3)
    h = C.c_void_p()
    p = (C.c_double * 2)(0.01, 1.0)
    with pytest.raises(Exception):
        capi.call("parm_integ_create", atoms._h, 42, p, 2, 0, C.byref(h))
    with pytest.raises(Exception):  # CollectionDamped: dt must be positive (collection.cpp:334-336)
        sim.CollectionDamped(sim.OriginBox(np.full(3, 5.0), 3), atoms, -1.0, 0.5)


@pytest.mark.parametrize("ndim", [2, 3])
def test_nlcg_packer_matches_oracle(oracle_built, ndim):
    """CollectionNLCG as pyparm/packmin.py drives it: compress bidisperse harmonic spheres towards P0. The secant
    loop branches on dot products, so device and CPU are compared over a stretch short enough that no branch sits
    within rounding of its threshold, then the GPU run is continued to convergence on its own."""
    from parm_b200 import sim
    w = W.packer_system(ndim=ndim, n=600, seed=4 + ndim)
    be = backends(oracle_built)[-1]
    box, atoms, inter, nl, collec = sim.from_workload(w)
    s = cpu_system(be, w, injected=True)
    for c in (collec, ):
        c.set_max_alpha(2.0)
        c.set_max_dx(10.0)
        c.set_max_step(1e-3)
    s.nlcg_set(3, 2.0)
    s.nlcg_set(5, 10.0)
    s.nlcg_set(6, 1e-3)
    collec.set_forces(True, True)
    s.nlcg_set_forces(True, True)
    for what in range(7):
        assert rel_err(collec._reduce(what), s.nlcg_reduce(what)) < 1e-10
    steps = 40
    for _ in range(steps):
        collec.timestep()
    s.timestep(steps)
    assert rel_err(box.box_shape(), s.get_box()) < 1e-9
    assert nl.which() == s.which()
    x = s.get_atoms()[0]
    assert rel_err_vec(atoms.peek("x") - w["x"], x - w["x"]) < 1e-7
    st, so = collec._state(), s.nlcg_get()
    assert rel_err(st[:8], so[:8]) < 1e-6
    for what in (4, 5, 6):
        assert rel_err(collec._reduce(what), s.nlcg_reduce(what)) < 1e-6
    # descend() and reset() (steepest-descent restart)
    collec.descend()
    s.nlcg_descend()
    collec.reset()
    s.nlcg_reset()
    assert rel_err(box.box_shape(), s.get_box()) < 1e-9
    assert rel_err(collec._state()[:8], s.nlcg_get()[:8]) < 1e-6
    # on its own: the packing approaches the goal pressure, the box shrinks
    V0 = box.V()
    collec.set_max_step(1e-2)
    collec.timestep(1500)
    assert box.V() < V0
    assert abs(collec.pressure() / collec.P0 - 1) < 0.2


def test_box_resize_keeps_the_pair_list(oracle_built):
    """OriginBox::resize_to does not touch the NeighborList (box.cpp:21-25): forces use the new box, the list is
    only rebuilt when the drift rule fires."""
    from parm_b200 import sim
    w = W.lj_lattice((8, 8, 8), seed=2)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    s = cpu_system(backends(oracle_built)[-1], w, injected=True)
    box.resize(0.97)
    s.set_box(np.asarray(w["L"]) * 0.97)
    collec.set_forces(True)
    s.set_forces(True)
    assert nl.which() == s.which() == 1
    assert rel_err_vec(atoms.peek("f"), s.get_atoms()[3]) < 1e-10
    collec.timestep(50)
    s.timestep(50)
    assert nl.which() == s.which()
    assert rel_err(collec.energy(), s.energy()) < 1e-9
