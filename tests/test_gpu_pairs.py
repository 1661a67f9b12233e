"""every NListed functor of sim.i:621-643 on the GPU against the oracle (the compiled reference when
present, else its C restatement): pair set bit-exact, forces / energy / virial / stress to 1e-10, contact and
overlap counts exact, and a short NVE trajectory -- with tabulated species and with per-atom continuous
parameters (pair constructor on the device), in 3-D and 2-D."""
import numpy as np
import pytest

from parm_b200 import workloads as W
from parity_util import backends, cpu_system, rel_err, rel_err_vec

pytestmark = pytest.mark.gpu
IDS = ["k%d%s" % c for c in W.FUNCTOR_CASES]


def check_static(w, be):
    from parm_b200 import sim
    box, atoms, inter, nl, collec = sim.from_workload(w)
    s = cpu_system(be, w, injected=True)
    a, b = nl.pairs()
    ra, rb = s.pairs()
    assert len(ra) > 0 and np.array_equal(a, ra) and np.array_equal(b, rb)
    atoms.reset_forces()
    p = inter.set_forces_get_pressure(box)
    f_ref, p_ref = s.forces_and_pressure()
    assert np.abs(f_ref).max() > 0
    assert rel_err_vec(atoms.peek("f"), f_ref) < 1e-10
    assert rel_err(p, p_ref) < 1e-10
    assert rel_err(inter.energy(box), s.inter_energy()) < 1e-10
    assert rel_err(inter.pressure(box), s.inter_pressure()) < 1e-10
    assert rel_err(inter.stress(box), s.inter_stress()) < 1e-10
    atoms.reset_forces()
    st = inter.set_forces_get_stress(box)
    assert rel_err(st, s.inter_stress()) < 1e-10
    assert rel_err_vec(atoms.peek("f"), f_ref) < 1e-10
    assert (inter.contacts(box), inter.overlaps(box)) == s.inter_contacts()
    return box, atoms, inter, nl, collec, s


@pytest.mark.parametrize("kind,variant", W.FUNCTOR_CASES, ids=IDS)
def test_functor_matches_oracle_3d(oracle_built, kind, variant):
    w = W.functor_system(kind, variant, ndim=3, n=3000, seed=40 + kind)
    for be in backends(oracle_built)[-1:]:
        box, atoms, inter, nl, collec, s = check_static(w, be)
        collec.set_forces(True)
        s.set_forces(True)
        collec.timestep(150)
        s.timestep(150)
        x, v, a, f = s.get_atoms()
        assert nl.which() == s.which()
        assert rel_err_vec(atoms.peek("x") - w["x"], x - w["x"]) < 1e-8
        assert rel_err_vec(atoms.peek("v"), v) < 1e-8
        assert rel_err(collec.energy(), s.energy()) < 1e-9
        assert rel_err(collec.pressure(), s.pressure()) < 1e-8


@pytest.mark.parametrize("kind,variant", W.FUNCTOR_CASES, ids=IDS)
def test_functor_continuous_parameters_2d(oracle_built, kind, variant):
    """> 32 distinct parameter tuples: the pair constructor runs per pair on the device (pairs.cuh mix_pair)."""
    w = W.functor_system(kind, variant, ndim=2, n=2500, seed=70 + kind, continuous=True)
    for be in backends(oracle_built)[-1:]:
        box, atoms, inter, nl, collec, s = check_static(w, be)
        collec.set_forces(True)
        s.set_forces(True)
        collec.timestep(60)
        s.timestep(60)
        x, v, a, f = s.get_atoms()
        assert rel_err_vec(atoms.peek("x") - w["x"], x - w["x"]) < 1e-8
        assert rel_err(collec.energy(), s.energy()) < 1e-9


def test_unknown_functor_and_bad_tables_fail_loudly():
    from parm_b200 import capi, sim
    w = W.functor_system(W.KIND_LJISH, n=64, seed=1)
    box = sim.OriginBox(w["L"], 3)
    atoms = sim.AtomVec(w["m"], ndim=3)
    nl = sim.NeighborList(box, atoms, 0.3)
    import ctypes as C
    h = C.c_void_p()
    with pytest.raises(Exception):
        capi.call("parm_inter_create", atoms._h, nl._h, 99, C.byref(h))
    inter = sim.LJish(atoms, nl)
    inter.add_many(w["params"], w["types"], None)  # LJishPair needs the epsilon table
    with pytest.raises(Exception):
        inter.energy(box)
    bad = w["eps_table"].copy()
    bad[0, 1] += 1.0  # asymmetric
    inter.add_many(w["params"], w["types"], bad)
    with pytest.raises(Exception):
        inter.energy(box)


@pytest.mark.parametrize("ndim", [3, 2])
def test_neighborlist_ignore(oracle_built, ndim):
    """NeighborList::ignore (trackers.hpp:190-193): pair set bit-exact against the oracle with bonded-style
    exclusions, the rebuild it forces, and the trajectory that follows."""
    from parm_b200 import sim
    from test_oracle_pin import chain_ignores
    n = 4000
    w = W.random_system(n, ndim, 2, seed=21 + ndim, ntypes=2)
    a, b = chain_ignores(n, np.random.default_rng(3))
    box, atoms, inter, nl, collec = sim.from_workload(w)
    s = cpu_system(backends(oracle_built)[-1], w, injected=True)
    n0 = nl.numpairs()
    w0 = nl.which()
    nl.ignore(a, b)
    s.ignore(a, b)
    assert nl.ignore_size() == s.ignore_size() > 0
    assert nl.update_list(False) and s.update_list(False)  # ignorechanged forces the rebuild (trackers.cpp:23)
    collec.set_forces(True)
    s.set_forces(True)
    assert nl.which() == w0 + 1 == s.which()
    pa, pb = nl.pairs()
    ra, rb = s.pairs()
    assert 0 < len(ra) < n0 and np.array_equal(pa, ra) and np.array_equal(pb, rb)
    assert rel_err_vec(atoms.peek("f"), s.get_atoms()[3]) < 1e-10
    assert rel_err(inter.energy(box), s.inter_energy()) < 1e-10
    collec.timestep(120)
    s.timestep(120)
    assert nl.which() == s.which() > w0 + 1
    pa, pb = nl.pairs()
    ra, rb = s.pairs()
    assert np.array_equal(pa, ra) and np.array_equal(pb, rb)
    assert rel_err_vec(atoms.peek("x") - w["x"], s.get_atoms()[0] - w["x"]) < 1e-8
    assert rel_err(collec.energy(), s.energy()) < 1e-9
    # AtomID form, one pair at a time
    nl.ignore(sim.AtomID(atoms, 10), sim.AtomID(atoms, 500))
    assert nl.ignore_size() == s.ignore_size() + 1


@pytest.mark.parametrize("nspecies", [2, 3])
def test_species_packed_into_list_entries(oracle_built, nspecies):
    """Long rows + 2..32 tabulated species: the neighbour's species id rides in the top bits of the row entries
    (csrc/nlist.cu k_pack_species; two species use the register path of the force kernel). Covers the state
    changes around it: parameters changed after the build (entries stale until the next rebuild), a second
    interaction sharing the list (it must mask the entries and gather its own species), ignore()."""
    from parm_b200 import sim
    w = W.lj_lattice((14, 14, 14), seed=8)  # ~110 neighbours per atom
    n = w["x"].shape[0]
    rng = np.random.default_rng(nspecies)
    types = rng.integers(0, nspecies, n).astype(np.uint32)
    tab = np.array([[1.0, 1.5, 0.7], [1.5, 0.5, 1.2], [0.7, 1.2, 0.9]])[:nspecies, :nspecies]  # Kob-Andersen-like
    w.update(types=types, eps_table=tab)
    w["params"][:, 1] = np.where(types == 1, 0.88, 1.0)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    s = cpu_system(backends(oracle_built)[-1], w, injected=True)

    def same():
        atoms.reset_forces()
        p = inter.set_forces_get_pressure(box)
        f_ref, p_ref = s.forces_and_pressure()
        assert rel_err_vec(atoms.peek("f"), f_ref) < 1e-10 and rel_err(p, p_ref) < 1e-10
        assert rel_err(inter.energy(box), s.inter_energy()) < 1e-10
        a, b = nl.pairs()
        ra, rb = s.pairs()
        assert np.array_equal(a, ra) and np.array_equal(b, rb)

    same()
    collec.set_forces(True)
    s.set_forces(True)
    collec.timestep(80)
    s.timestep(80)
    assert nl.which() == s.which() > 1
    assert rel_err_vec(atoms.peek("x") - w["x"], s.get_atoms()[0] - w["x"]) < 1e-8
    same()
    # a second interaction on the same list: soft repulsion with its own species (sigma classes)
    p2 = np.zeros((n, 3))
    p2[:, 0] = 2.0
    p2[:, 1] = np.where(rng.random(n) < 0.5, 0.9, 1.05)
    p2[:, 2] = 2.5
    rep = sim.Repulsion(atoms, nl)
    rep.add_many(p2)
    k2 = s.add_interaction(W.KIND_REPULSION, 0.0, p2, share_nl=0)
    atoms.reset_forces()
    rep.set_forces(box)
    s.api["reset_forces"](s.h)
    s.api["inter_set_forces"](s.h, k2)
    assert rel_err_vec(atoms.peek("f"), s.get_atoms()[3]) < 1e-10
    assert rel_err(rep.energy(box), s.inter_energy(k2)) < 1e-10
    same()
